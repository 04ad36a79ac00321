"""Full-size checks (BASELINE.json config 2: one ~2 M-voxel tile) through size-independent properties, where the oracle
would take minutes: partition / sortedness of the voxeliser, symmetry of the 3^3 rulebook, parent/child consistency of the
strided maps, linearity of the convolution, idempotence of the overlap merge, permutation invariance of the clustering and
a brute-force spot check of the kNN vote.  Integer properties are exact."""
import numpy as np
import pytest
import torch

from treelearn_b200 import TreeLearn, _lib, pipeline, sparse, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def big():
    batch = synth.make_batch([synth.workload('cfg2_2M')])
    dev = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
    vf, vc, keys, v2p = sparse.voxelize(dev['coords'], dev['input_feats'], dev['batch_ids'], 1, 0.1, False, True, 3)
    levels = sparse.build_levels(keys, vc, [1000, 1000, 1000], 7)
    return dev, vf, vc, keys, v2p, levels


def test_voxelize_partition_and_order_at_full_size(big):
    dev, vf, vc, keys, v2p, levels = big
    n, m = dev['coords'].shape[0], vc.shape[0]
    assert m > 1_500_000
    assert bool((keys[1:] > keys[:-1]).all())                                  # Morton keys strictly ascending: unique voxels
    assert int(v2p.min()) == 0 and int(v2p.max()) == m - 1
    assert int(torch.unique(v2p).numel()) == m                                 # every voxel owns at least one point
    # every point lies inside the voxel it was assigned to: floor((p - min) / 0.1) == voxel coordinate (fp32 rule)
    mn = dev['coords'].min(dim=0).values
    vs = torch.tensor(0.1, dtype=torch.float32, device='cuda')                  # IEEE fp32 division, not a reciprocal multiply
    idx = torch.floor((dev['coords'] - mn) / vs).to(torch.int32)
    assert bool((idx == vc[v2p][:, 1:]).all())
    # mean pooled verticality: voxels with a single point carry that point's feature exactly
    counts = torch.bincount(v2p, minlength=m)
    single = counts[v2p] == 1
    assert torch.equal(vf[v2p[single], 0], dev['input_feats'][single, 0])


def test_rulebook_symmetry_and_geometry_at_full_size(big):
    _, _, vc, _, _, levels = big
    for lv in levels[:4]:
        n = lv.n
        nbr = lv.nbr[:, :n]
        rows = torch.arange(n, device=nbr.device, dtype=torch.int32)
        assert torch.equal(nbr[13], rows)                                      # centre tap = identity
        for k in (0, 4, 9, 12, 22, 26):
            j = nbr[k]
            ok = j >= 0
            jj = j[ok].long()
            assert torch.equal(nbr[26 - k][jj], rows[ok])                      # nbr_k(i) = j  <=>  nbr_{26-k}(j) = i
            d = torch.tensor([k // 9 - 1, (k // 3) % 3 - 1, k % 3 - 1], device=nbr.device, dtype=torch.int32)
            assert bool((lv.coords[jj][:, 1:] == lv.coords[ok][:, 1:] + d).all())   # the neighbour really sits at p + delta
        # tile masks: bit k set iff some row of the 128-row tile has offset k
        k = 5
        has = (nbr[k] >= 0)
        pad = (-n) % 128
        has = torch.cat([has, torch.zeros(pad, dtype=torch.bool, device=has.device)]).view(-1, 128).any(dim=1)
        assert torch.equal(((lv.nbr_mask[:has.numel()] >> k) & 1).bool(), has)


def test_strided_maps_are_consistent_at_full_size(big):
    _, _, _, _, _, levels = big
    fine, coarse = levels[0], levels[1]
    up = fine.up_index[:, :fine.n]
    par = up.max(dim=0).values                                                 # exactly one kappa per fine row
    assert bool((par >= 0).all()) and bool(((up >= 0).sum(dim=0) == 1).all())
    assert bool((coarse.coords[par.long()][:, 1:] == fine.coords[:, 1:] // 2).all())
    down = fine.down_index[:, :coarse.n]
    ok = down >= 0
    assert int(ok.sum()) == fine.n                                             # every fine row feeds exactly one coarse row
    q = torch.arange(coarse.n, device=down.device).expand(8, -1)[ok]
    assert torch.equal(par[down[ok].long()].long(), q)


def test_conv_linearity_at_full_size(big):
    _, _, _, _, _, levels = big
    lv = levels[1]                                                            # ~0.6 M voxels, 64 channels
    g = torch.Generator(device='cuda').manual_seed(0)
    x = torch.randn((lv.n, 64), device='cuda', generator=g)
    y = torch.randn((lv.n, 64), device='cuda', generator=g)
    w = torch.randn((27, 64, 64), device='cuda', generator=g) / 40
    wp = w.contiguous()                                                        # fp32 SIMT layout [K, C_in, C_out]
    f = lambda t: sparse.conv([sparse.Seg(t, wp, lv.nbr, lv.nbr_mask)], lv.n, 64, _lib.MODE_FP32, raw=True)  # noqa: E731
    lhs = f(2.0 * x - 0.5 * y)
    rhs = 2.0 * f(x) - 0.5 * f(y)
    assert torch.allclose(lhs, rhs, atol=2e-4, rtol=1e-4)
    # tcgen05 fp16-operand path against the fp32 path on fp16-representable data
    xh = x.half()
    wh = (w.half()).float()
    ref = sparse.conv([sparse.Seg(xh.float(), wh.contiguous(), lv.nbr, lv.nbr_mask)], lv.n, 64, _lib.MODE_FP32, raw=True)
    wk = wh.permute(0, 2, 1)                                                   # [K, C_out, C_in]
    pw = sparse.pack_weight(wk, 1) if sparse.USE_TS else sparse.pack_weight_tc(wk, True)
    to_k, from_k = (sparse.to_p, sparse.from_p) if sparse.USE_TS else ((lambda a: a), (lambda a: a))   # kernel layout
    tc = from_k(sparse.conv([sparse.Seg(to_k(xh).contiguous(), pw, lv.nbr, lv.nbr_mask)], lv.n, 64, _lib.MODE_F16, raw=True))
    assert torch.allclose(tc, ref, atol=3e-4, rtol=2e-4)                       # only the accumulation order differs


def test_default_model_f16_close_to_fp32_at_full_size(big):
    dev = big[0]
    sd = None
    outs = {}
    for mode in ('fp32', 'f16'):
        torch.manual_seed(0)
        net = synth.randomize_bn_stats(TreeLearn(use_feats=False, use_coords=False, spatial_shape=[1000, 1000, 1000], mode=mode))
        if sd is None:
            sd = net.state_dict()
        net.load_state_dict(sd)
        net = net.cuda().eval()
        with torch.no_grad():
            outs[mode] = net({k: dev[k] for k in ('coords', 'input_feats', 'batch_ids', 'batch_size')}, return_loss=False)
    err = (outs['f16']['offset_predictions'] - outs['fp32']['offset_predictions']).abs().max().item()
    assert err < 1e-3, err                                                     # north_star tolerance, 2.1 M points


def test_merge_idempotent_and_cluster_permutation_invariant(big):
    dev = big[0]
    g = torch.Generator(device='cuda').manual_seed(1)
    n = 1_000_000
    xyz = (torch.rand((n, 3), device='cuda', generator=g) * torch.tensor([30.0, 30.0, 20.0], device='cuda'))
    xyz = torch.round(xyz * 20) / 20                                           # 5 cm lattice -> plenty of duplicates
    vals = torch.rand((n, 6), device='cuda', generator=g)
    c1, v1 = pipeline.ensemble_cuda(xyz, vals)
    c2, v2 = pipeline.ensemble_cuda(c1, v1)
    assert torch.equal(c1, c2) and torch.allclose(v1, v2, atol=1e-6)           # merging merged rows changes nothing
    key = (c1 * 100).round().long()
    lin = (key[:, 0] * 10_000 + key[:, 1]) * 10_000 + key[:, 2]
    assert bool((lin[1:] > lin[:-1]).all())                                    # sorted by (x, y, z), unique
    # clustering: the partition does not depend on the order of the points
    pts = (dev['coords'][:400_000, :2] + 0.03 * torch.randn((400_000, 2), device='cuda', generator=g)).contiguous()
    lab, ncl = pipeline.group_dbscan_cuda(pts, 0.15, 50, -1, 1)
    perm = torch.randperm(pts.shape[0], device='cuda', generator=g)
    lab_p, ncl_p = pipeline.group_dbscan_cuda(pts[perm].contiguous(), 0.15, 50, -1, 1)
    assert ncl == ncl_p
    a, b = lab[perm], lab_p
    assert torch.equal(a == -1, b == -1)
    pairs = torch.unique(torch.stack([a[a > 0], b[b > 0]], dim=1), dim=0)
    assert pairs.shape[0] == ncl                                               # one-to-one label correspondence


def test_knn_vote_spot_check_against_brute_force(big):
    g = torch.Generator(device='cuda').manual_seed(2)
    ref = torch.rand((300_000, 3), device='cuda', generator=g) * 40
    lab = torch.randint(1, 60, (300_000,), device='cuda', generator=g)
    qry = torch.rand((200_000, 3), device='cuda', generator=g) * 40
    out = pipeline.knn_vote_cuda(ref, lab, qry, 5)
    sel = torch.arange(0, 200_000, 397, device='cuda')
    d = torch.cdist(qry[sel].double(), ref.double())
    nn = d.topk(5, dim=1, largest=False).indices
    votes = lab[nn]
    cnt = (votes[:, :, None] == votes[:, None, :]).sum(-1)
    want = torch.where(cnt == cnt.max(dim=1, keepdim=True).values, votes, torch.full_like(votes, 1 << 40)).min(dim=1).values
    # (most frequent label among the 5 nearest, ties -> smallest label)
    assert torch.equal(out[sel], want)
