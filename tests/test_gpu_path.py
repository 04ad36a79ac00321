"""GPU parity tests (run on the B200 box): every CUDA op called through the C-ABI against the oracle on the
same seeded inputs.  Integer / index work must be bit-exact; floating point within the stated tolerance."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import cluster_ref, model_ref, spconv_ref as sp
from treelearn_b200 import TreeLearn, pipeline, sparse, synth, _lib

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')
FP32_TOL = dict(atol=2e-4, rtol=1e-4)     # fp32 SIMT path: only summation order differs from the oracle


def _tile(name='tiny'):
    return synth.make_batch([synth.workload(name)])


def _key_rows(idx):
    """Map (b,x,y,z) rows -> dict for order-free comparison."""
    return {tuple(r): i for i, r in enumerate(np.asarray(idx).tolist())}


def _voxelize_both(batch, use_feats=True):
    dev = 'cuda'
    vf, vc, keys, v2p = sparse.voxelize(batch['coords'].to(dev), batch['input_feats'].to(dev),
                                        batch['batch_ids'].to(dev), batch['batch_size'], 0.1, False, use_feats, 3)
    ovf, ovi, ov2p, oshape = model_ref.voxelize_ref(batch['coords'], batch['input_feats'], batch['batch_ids'],
                                                    batch['batch_size'], 0.1, False, use_feats, 3)
    return (vf, vc, keys, v2p), (ovf, ovi, ov2p, oshape)


def test_voxelize_bit_exact_partition_and_mean():
    tiles = [synth.synth_forest(edge=5.0, n_trees=2, seed=s, ground_density=200.0) for s in (1, 2)]
    # un-rounded duplicates so that voxels hold several points (mean-pool of the first <=3 is exercised)
    for t in tiles:
        t['coords'] = np.concatenate([t['coords'], t['coords'][::3] + np.float32(0.013)])
        t['feat'] = np.concatenate([t['feat'], t['feat'][::3] * np.float32(0.5)])
        t['inst'] = np.concatenate([t['inst'], t['inst'][::3]])
        t['base'] = np.concatenate([t['base'], t['base'][::3]])
    batch = synth.make_batch(tiles)
    (vf, vc, keys, v2p), (ovf, ovi, ov2p, _) = _voxelize_both(batch)
    vc, v2p, vf = vc.cpu().numpy(), v2p.cpu().numpy(), vf.cpu().numpy()
    assert len(vc) == len(ovi)
    assert np.all(np.diff(keys.cpu().numpy()) > 0), 'voxel keys must be strictly ascending (Morton order)'
    mine, theirs = _key_rows(vc), _key_rows(ovi.numpy())
    assert set(mine) == set(theirs)                                   # same voxel set, bit exact
    # same point -> voxel partition: compare the voxel coordinates each point lands in
    assert np.array_equal(vc[v2p], ovi.numpy()[ov2p.numpy()])
    perm = np.array([theirs[tuple(r)] for r in vc.tolist()])
    assert np.array_equal(vf, ovf.numpy()[perm]), 'mean-pooled features must be bit exact'


def test_rulebooks_bit_exact():
    batch = synth.make_batch([synth.synth_forest(edge=5.0, n_trees=2, seed=5, ground_density=200.0),
                              synth.synth_forest(edge=4.0, n_trees=1, seed=6, ground_density=200.0)])
    (vf, vc, keys, v2p), _ = _voxelize_both(batch)
    shape = [40, 41, 130]     # deliberately tight and odd so the bounds rule and the odd-edge rule both bite
    levels = sparse.build_levels(keys, vc, shape, 4)
    idx = vc.cpu().numpy()
    for l, lv in enumerate(levels):
        coords = lv.coords.cpu().numpy()
        n = lv.n
        assert n == len(coords)
        ref_nbr = sp.subm_neighbour_table(coords, lv.shape, 3)          # rows in MY order => directly comparable
        mine = lv.nbr.cpu().numpy()[:, :n].astype(np.int64)
        # voxels outside spatial_shape: spconv semantics undefined; this library keeps the centre tap only
        inside = np.all(coords[:, 1:] < np.asarray(lv.shape)[None, :], axis=1)
        assert np.array_equal(mine[:, inside], ref_nbr[:, inside]), f'subm rulebook level {l}'
        assert np.all(lv.nbr.cpu().numpy()[:, n:] == -1)
        tm = lv.nbr_mask.cpu().numpy().astype(np.uint32)
        for t in range(len(tm)):
            rows = mine[:, t * 128:(t + 1) * 128]
            expect = sum(1 << k for k in range(27) if (rows[k] >= 0).any())
            assert int(tm[t]) == expect
        if l + 1 < len(levels):
            out_idx, out_shape, in_row, kappa, out_row = sp.strided_pairs(coords, lv.shape)
            nxt = levels[l + 1]
            assert nxt.shape == out_shape and nxt.n == len(out_idx)
            ccoords = nxt.coords.cpu().numpy()
            where = _key_rows(ccoords)                      # coarse row order is implementation-defined (Morton here)
            assert set(where) == set(_key_rows(out_idx)), 'coarse voxel set'
            assert np.all(np.diff(nxt.keys.cpu().numpy()) > 0)
            to_mine = np.array([where[tuple(r)] for r in out_idx.tolist()])
            out_row = to_mine[out_row]
            down = lv.down_index.cpu().numpy()
            up = lv.up_index.cpu().numpy()
            exp_down = -np.ones_like(down)
            exp_up = -np.ones_like(up)
            exp_down[kappa, out_row] = in_row
            exp_up[kappa, in_row] = out_row
            assert np.array_equal(down, exp_down) and np.array_equal(up, exp_up)


def test_reach_zero_error_contract():
    batch = _tile()
    (vf, vc, keys, v2p), _ = _voxelize_both(batch)
    with pytest.raises(ValueError, match='reach zero!!!'):
        sparse.build_levels(keys, vc, [70, 70, 210], 8)


@pytest.mark.parametrize('ci,co', [(4, 32), (32, 32), (64, 32), (96, 96), (224, 224), (8, 24)])
def test_subm_conv_layer_parity(ci, co):
    batch = synth.make_batch([synth.synth_forest(edge=4.0, n_trees=2, seed=9, ground_density=150.0)])
    (vf, vc, keys, v2p), _ = _voxelize_both(batch)
    lv = sparse.build_levels(keys, vc, [500, 500, 1000], 1)[0]
    g = torch.Generator().manual_seed(ci * 1000 + co)
    x = torch.randn((lv.n, ci), generator=g)
    w = torch.randn((co, 3, 3, 3, ci), generator=g) / (27 * ci) ** 0.5
    res = torch.randn((lv.n, co), generator=g)
    s, t = torch.rand(co, generator=g) + 0.5, torch.randn(co, generator=g)
    ref = model_ref._subm(x, sp.subm_neighbour_table(vc.cpu().numpy(), [500, 500, 1000]), w) + res
    wp = w.reshape(co, 27, ci).permute(1, 2, 0).contiguous().cuda()
    raw, act = sparse.conv([sparse.Seg(x.cuda(), wp, lv.nbr, lv.nbr_mask)], lv.n, co, _lib.MODE_FP32,
                           residual=res.cuda(), raw=True, act1=(s.cuda(), t.cuda()))
    assert torch.allclose(raw.cpu(), ref, **FP32_TOL)
    assert torch.allclose(act.cpu(), torch.relu(ref * s + t), **FP32_TOL)


def test_strided_and_inverse_conv_parity():
    batch = synth.make_batch([synth.synth_forest(edge=4.0, n_trees=2, seed=10, ground_density=150.0)])
    (vf, vc, keys, v2p), _ = _voxelize_both(batch)
    lv, nx = sparse.build_levels(keys, vc, [500, 500, 1000], 2)
    g = torch.Generator().manual_seed(3)
    x = torch.randn((lv.n, 32), generator=g)
    wd = torch.randn((64, 2, 2, 2, 32), generator=g) / 16
    wu = torch.randn((32, 2, 2, 2, 64), generator=g) / 16
    out_idx, out_shape, in_row, kappa, out_row = sp.strided_pairs(vc.cpu().numpy(), [500, 500, 1000])
    where = _key_rows(nx.coords.cpu().numpy())          # coarse rows: my (Morton) order vs the oracle's
    out_row = np.array([where[tuple(r)] for r in out_idx.tolist()])[out_row]
    ref_d = model_ref._pairs_conv(x, wd, in_row, kappa, out_row, len(out_idx))
    ref_u = model_ref._pairs_conv(ref_d, wu, out_row, kappa, in_row, lv.n)
    d = sparse.conv([sparse.Seg(x.cuda(), wd.reshape(64, 8, 32).permute(1, 2, 0).contiguous().cuda(), lv.down_index,
                                lv.down_mask)], nx.n, 64, _lib.MODE_FP32, raw=True)
    u = sparse.conv([sparse.Seg(d, wu.reshape(32, 8, 64).permute(1, 2, 0).contiguous().cuda(), lv.up_index,
                                lv.up_mask)], lv.n, 32, _lib.MODE_FP32, raw=True)
    assert torch.allclose(d.cpu(), ref_d, **FP32_TOL)
    assert torch.allclose(u.cpu(), ref_u, **FP32_TOL)


def _load_model_fixture():
    g = np.load(os.path.join(GOLD, 'model_small.npz'))
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith('sd:')}
    batch = {k[6:]: torch.from_numpy(g[k]) for k in g.files if k.startswith('batch:')}
    batch['batch_size'] = int(batch['batch_size'])
    out = {k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith('out:')}
    return g, sd, batch, out


def test_model_matches_reference_code_golden():
    """Golden vectors produced by the reference's own model code (tests/golden/make_golden.py)."""
    g, sd, batch, out = _load_model_fixture()
    net = TreeLearn(channels=8, num_blocks=3, use_feats=True, use_coords=False, spatial_shape=[500, 500, 1000])
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    with torch.no_grad():
        mine = net(batch, return_loss=False)
        loss, ld = net(batch, return_loss=True)
    for k, v in out.items():
        assert mine[k].is_cuda
        assert torch.allclose(mine[k].cpu(), v, **FP32_TOL), k
    assert abs(loss.item() - float(g['loss'])) < 1e-3
    assert abs(ld['semantic_loss'].item() - float(g['semantic_loss'])) < 1e-3
    assert abs(ld['offset_loss'].item() - float(g['offset_loss'])) < 1e-3


def test_default_model_vs_oracle_offsets_within_1e3():
    """Default 7-level / 32-channel network on a synthetic tile: per-point offsets within 1e-3 (north_star)."""
    batch = _tile('tiny')
    sd = model_ref.make_state_dict(channels=32, num_blocks=7, seed=0)
    net = TreeLearn(use_feats=False, use_coords=False, spatial_shape=[500, 500, 1000])
    net.load_state_dict(sd)
    net = net.cuda().eval()
    with torch.no_grad():
        mine = net(batch, return_loss=False)
        ref = model_ref.forward_ref(sd, batch, spatial_shape=[500, 500, 1000])
    assert (mine['offset_predictions'].cpu() - ref['offset_predictions']).abs().max() < 1e-3
    assert (mine['semantic_prediction_logits'].cpu() - ref['semantic_prediction_logits']).abs().max() < 1e-3
    assert torch.allclose(mine['backbone_feats'].cpu(), ref['backbone_feats'], atol=1e-3, rtol=1e-3)


def test_spatial_shape_none_and_batch_of_two():
    tiles = [synth.synth_forest(edge=4.0, n_trees=1, seed=s, ground_density=120.0) for s in (31, 32)]
    batch = synth.make_batch(tiles)
    sd = model_ref.make_state_dict(channels=8, num_blocks=3, seed=2)
    net = TreeLearn(channels=8, num_blocks=3, use_feats=True, use_coords=True, spatial_shape=None)
    net.load_state_dict(sd)
    net = net.cuda().eval()
    with torch.no_grad():
        mine = net(batch, return_loss=False)
        ref = model_ref.forward_ref(sd, batch, use_coords=True, use_feats=True, spatial_shape=None)
    for k in ref:
        assert torch.allclose(mine[k].cpu(), ref[k], atol=5e-4, rtol=1e-3), k


# ---- merge / clustering / kNN --------------------------------------------------------------------
def _cluster_fixture():
    return np.load(os.path.join(GOLD, 'cluster_small.npz'))


NAMES = ['coords', 'semantic_scores', 'semantic_labels', 'offset_predictions', 'offset_labels', 'instance_labels',
         'feats', 'input_feats']


def test_ensemble_matches_reference_golden():
    g = _cluster_fixture()
    out = pipeline.ensemble(**{n: g['ens_in:' + n] for n in NAMES})
    for n, a in zip(NAMES, out):
        ref = g['ens_out:' + n]
        assert a.dtype == ref.dtype and a.shape == ref.shape, n
        if a.dtype == np.int64:
            assert np.array_equal(a, ref), n
        else:
            assert np.allclose(a, ref, rtol=1e-5, atol=1e-6), n
    assert np.array_equal(out[0], g['ens_out:coords']), 'merged coordinates and their order are bit exact'


def test_get_instances_and_knn_match_reference_golden():
    g = _cluster_fixture()
    cfg = SimpleNamespace(tree_conf_thresh=0.5, tau_vert=0.6, tau_off=4, tau_group=0.15, tau_min=50, use_hdbscan=False)
    coords, logits, offs, vert = g['ens_out:coords'], g['ens_out:semantic_scores'], g['ens_out:offset_predictions'], \
        g['ens_out:input_feats'][:, -1]
    inst = pipeline.get_instances(coords, offs, logits, cfg, vert, 0, 0, -1, 1)
    assert inst.dtype == np.int64 and np.array_equal(inst, g['instances'])
    tm = inst != 0
    assigned = pipeline.assign_remaining_points_nearest_neighbor(coords[tm] + offs[tm], inst[tm], -1)
    assert np.array_equal(assigned, g['assigned'])


def test_group_dbscan_label_ids_match_sklearn_golden():
    g = _cluster_fixture()
    assert np.array_equal(pipeline.group_dbscan(g['p2'], 0.15, 20, -1, 1), g['p2_group'])
    raw = pipeline.group_dbscan(g['p2'], 0.15, 0, -1, 0)          # no size filter: raw DBSCAN numbering
    assert np.array_equal(raw, g['p2_raw'])


@pytest.mark.parametrize('n,seed', [(1, 0), (2, 1), (777, 2), (20000, 3)])
def test_cluster_and_knn_vs_oracle_random(n, seed):
    rng = np.random.default_rng(seed)
    pts = (rng.normal(size=(n, 2)) * 0.6).astype(np.float32)
    pts[: n // 3] = np.round(pts[: n // 3], 1)                   # exact duplicates and exact-distance ties
    assert np.array_equal(pipeline.group_dbscan(pts, 0.15, 3, -1, 1), cluster_ref.group_dbscan_ref(pts, 0.15, 3, -1, 1))
    if n >= 100:
        xyz = (rng.normal(size=(n, 3)) * np.array([2.0, 2.0, 6.0])).astype(np.float32)
        pred = rng.integers(0, 6, n) - 1
        pred[:5] = np.arange(5)
        assert np.array_equal(pipeline.assign_remaining_points_nearest_neighbor(xyz, pred, -1),
                              cluster_ref.assign_remaining_ref(xyz, pred, -1))


# ---- tcgen05 / TF32 path ---------------------------------------------------------------------------
from treelearn_b200.model import _round_tf32  # noqa: E402

TF32_EXACT_TOL = dict(atol=3e-4, rtol=2e-4)    # operands pre-rounded to TF32: only the accumulation order differs


def _tc_weight(w, co, k, ci):
    return sparse.pack_weight_tc(w.reshape(co, k, ci).permute(1, 0, 2).cuda(), False)


@pytest.mark.parametrize('ci,co', [(32, 32), (64, 32), (32, 64), (96, 96), (128, 160), (224, 224)])
def test_tc_subm_conv_parity(ci, co):
    batch = synth.make_batch([synth.synth_forest(edge=5.0, n_trees=2, seed=9, ground_density=200.0)])
    (vf, vc, keys, v2p), _ = _voxelize_both(batch)
    lv = sparse.build_levels(keys, vc, [500, 500, 1000], 1)[0]
    g = torch.Generator().manual_seed(ci * 1000 + co)
    x = _round_tf32(torch.randn((lv.n, ci), generator=g))
    w = _round_tf32(torch.randn((co, 3, 3, 3, ci), generator=g) / (27 * ci) ** 0.5)
    res = torch.randn((lv.n, co), generator=g)
    s, t = torch.rand(co, generator=g) + 0.5, torch.randn(co, generator=g)
    ref = model_ref._subm(x, sp.subm_neighbour_table(vc.cpu().numpy(), [500, 500, 1000]), w) + res
    raw, act, act2 = sparse.conv([sparse.Seg(x.cuda(), _tc_weight(w, co, 27, ci), lv.nbr, lv.nbr_mask)], lv.n, co,
                                 _lib.MODE_TF32, residual=res.cuda(), raw=True, act1=(s.cuda(), t.cuda()),
                                 act2=(t.cuda().abs() + 0.1, s.cuda()))
    assert torch.allclose(raw.cpu(), ref, **TF32_EXACT_TOL)
    assert torch.allclose(act.cpu(), torch.relu(ref * s + t), atol=2e-3, rtol=1e-3)      # + TF32 rounding of the store
    assert torch.allclose(act2.cpu(), torch.relu(ref * (t.abs() + 0.1) + s), atol=2e-3, rtol=1e-3)


def test_tc_multi_segment_strided_inverse_parity():
    batch = synth.make_batch([synth.synth_forest(edge=5.0, n_trees=2, seed=10, ground_density=200.0)])
    (vf, vc, keys, v2p), _ = _voxelize_both(batch)
    lv, nx = sparse.build_levels(keys, vc, [500, 500, 1000], 2)
    g = torch.Generator().manual_seed(3)
    rt = lambda *shape: _round_tf32(torch.randn(shape, generator=g))   # noqa: E731
    x, wd, wu = rt(lv.n, 32), rt(64, 2, 2, 2, 32) / 16, rt(32, 2, 2, 2, 64) / 16
    out_idx, out_shape, in_row, kappa, out_row = sp.strided_pairs(vc.cpu().numpy(), [500, 500, 1000])
    where = _key_rows(nx.coords.cpu().numpy())
    out_row = np.array([where[tuple(r)] for r in out_idx.tolist()])[out_row]
    ref_d = model_ref._pairs_conv(x, wd, in_row, kappa, out_row, nx.n)
    d = sparse.conv([sparse.Seg(x.cuda(), _tc_weight(wd, 64, 8, 32), lv.down_index, lv.down_mask)], nx.n, 64,
                    _lib.MODE_TF32, raw=True)
    assert torch.allclose(d.cpu(), ref_d, **TF32_EXACT_TOL)
    d_r = _round_tf32(ref_d)
    ref_u = model_ref._pairs_conv(d_r, wu, out_row, kappa, in_row, lv.n)
    u = sparse.conv([sparse.Seg(d_r.cuda(), _tc_weight(wu, 32, 8, 64), lv.up_index, lv.up_mask)], lv.n, 32,
                    _lib.MODE_TF32, raw=True)
    assert torch.allclose(u.cpu(), ref_u, **TF32_EXACT_TOL)
    # three segments: 3^3 conv + two identity (1x1) segments == blocks_tail.block0 second conv
    h, z, e = rt(lv.n, 32), rt(lv.n, 32), rt(lv.n, 32)
    w3, wz, we = rt(32, 3, 3, 3, 32) / 32, rt(32, 1, 1, 1, 32) / 8, rt(32, 1, 1, 1, 32) / 8   # powers of two keep TF32 exactness
    nbr = sp.subm_neighbour_table(vc.cpu().numpy(), [500, 500, 1000])
    ref = model_ref._subm(h, nbr, w3) + z @ wz.reshape(32, 32).T + e @ we.reshape(32, 32).T
    out = sparse.conv([sparse.Seg(h.cuda(), _tc_weight(w3, 32, 27, 32), lv.nbr, lv.nbr_mask),
                       sparse.Seg(z.cuda(), _tc_weight(wz, 32, 1, 32)), sparse.Seg(e.cuda(), _tc_weight(we, 32, 1, 32))],
                      lv.n, 32, _lib.MODE_TF32, raw=True)
    assert torch.allclose(out.cpu(), ref, **TF32_EXACT_TOL)


def test_tc_default_model_offsets_within_1e3_of_fp32_oracle():
    """north_star tolerance: per-point offsets within 1e-3 of the (fp32) reference restatement, TF32 tensor cores."""
    batch = _tile('tiny')
    sd = model_ref.make_state_dict(channels=32, num_blocks=7, seed=0)
    net = TreeLearn(use_feats=False, use_coords=False, spatial_shape=[500, 500, 1000], mode='tf32')
    net.load_state_dict(sd)
    net = net.cuda().eval()
    with torch.no_grad():
        mine = net(batch, return_loss=False)
        ref = model_ref.forward_ref(sd, batch, spatial_shape=[500, 500, 1000])
    err = (mine['offset_predictions'].cpu() - ref['offset_predictions']).abs().max().item()
    print('tf32 offset max err', err)
    assert err < 1e-3
    assert (mine['semantic_prediction_logits'].cpu() - ref['semantic_prediction_logits']).abs().max() < 2e-3


def test_tc_matches_fp32_path_on_larger_tile():
    """No split-K at level 0 here (> 2 waves of tiles): the fused-epilogue tcgen05 path vs the fp32 SIMT path."""
    batch = _tile('small')
    sd = model_ref.make_state_dict(channels=32, num_blocks=7, seed=1)
    outs = {}
    for mode in ('fp32', 'tf32'):
        net = TreeLearn(use_feats=False, use_coords=False, spatial_shape=[500, 500, 1000], mode=mode)
        net.load_state_dict(sd)
        net = net.cuda().eval()
        with torch.no_grad():
            outs[mode] = net(batch, return_loss=False)
    err = (outs['tf32']['offset_predictions'] - outs['fp32']['offset_predictions']).abs().max().item()
    print('tf32 vs fp32 offsets', err)
    assert err < 1e-3
    assert (outs['tf32']['semantic_prediction_logits'] - outs['fp32']['semantic_prediction_logits']).abs().max() < 2e-3


# ---- tcgen05 / fp16-operand path -------------------------------------------------------------------
@pytest.mark.parametrize('ci,co', [(32, 32), (64, 32), (96, 96), (224, 224)])
def test_f16_subm_conv_parity(ci, co):
    batch = synth.make_batch([synth.synth_forest(edge=5.0, n_trees=2, seed=9, ground_density=200.0)])
    (vf, vc, keys, v2p), _ = _voxelize_both(batch)
    lv = sparse.build_levels(keys, vc, [500, 500, 1000], 1)[0]
    g = torch.Generator().manual_seed(ci * 1000 + co + 7)
    x = torch.randn((lv.n, ci), generator=g).half()
    w = (torch.randn((co, 3, 3, 3, ci), generator=g) / (27 * ci) ** 0.5).half()
    res = torch.randn((lv.n, co), generator=g)
    s, t = torch.rand(co, generator=g) + 0.5, torch.randn(co, generator=g)
    ref = model_ref._subm(x.float(), sp.subm_neighbour_table(vc.cpu().numpy(), [500, 500, 1000]), w.float()) + res
    wk = w.reshape(co, 27, ci).permute(1, 0, 2).cuda()
    wp = sparse.pack_weight(wk.float(), 1) if sparse.USE_TS else sparse.pack_weight_tc(wk, True)
    to_k, from_k = (sparse.to_p, sparse.from_p) if sparse.USE_TS else ((lambda a: a), (lambda a: a))   # kernel layout
    raw, act = sparse.conv([sparse.Seg(to_k(x.cuda()), wp, lv.nbr, lv.nbr_mask)], lv.n, co, _lib.MODE_F16,
                           residual=to_k(res.cuda()), raw=True, act1=(s.cuda(), t.cuda()))
    assert raw.dtype == torch.float32 and act.dtype == torch.float16
    raw, act = from_k(raw), from_k(act)
    assert torch.allclose(raw.cpu(), ref, **TF32_EXACT_TOL)
    assert torch.allclose(act.cpu().float(), torch.relu(ref * s + t), atol=3e-3, rtol=2e-3)   # + fp16 rounding of the store


def test_f16_default_model_offsets_within_1e3_of_fp32_oracle():
    batch = _tile('tiny')
    sd = model_ref.make_state_dict(channels=32, num_blocks=7, seed=0)
    net = TreeLearn(use_feats=False, use_coords=False, spatial_shape=[500, 500, 1000], mode='f16')
    net.load_state_dict(sd)
    net = net.cuda().eval()
    with torch.no_grad():
        mine = net(batch, return_loss=False)
        ref = model_ref.forward_ref(sd, batch, spatial_shape=[500, 500, 1000])
    err = (mine['offset_predictions'].cpu() - ref['offset_predictions']).abs().max().item()
    print('f16 offset max err', err)
    assert err < 1e-3
    assert (mine['semantic_prediction_logits'].cpu() - ref['semantic_prediction_logits']).abs().max() < 2e-3
    assert mine['backbone_feats'].dtype == torch.float32


def test_f16_matches_fp32_path_on_larger_tile():
    batch = _tile('small')
    sd = model_ref.make_state_dict(channels=32, num_blocks=7, seed=1)
    outs = {}
    for mode in ('fp32', 'f16'):
        net = TreeLearn(use_feats=False, use_coords=False, spatial_shape=[500, 500, 1000], mode=mode)
        net.load_state_dict(sd)
        net = net.cuda().eval()
        with torch.no_grad():
            outs[mode] = net(batch, return_loss=False)
    err = (outs['f16']['offset_predictions'] - outs['fp32']['offset_predictions']).abs().max().item()
    print('f16 vs fp32 offsets', err)
    assert err < 1e-3


def test_cluster_dense_blobs_vs_oracle():
    """Trained-model regime: every point of a tree lands on its base => thousands of points inside a few eps cells."""
    rng = np.random.default_rng(11)
    blobs = [rng.normal(size=(3000, 2)) * 0.08 + c for c in ([0, 0], [0.9, 0.1], [5, 5])]
    bridge = np.stack([np.linspace(0.25, 0.65, 4), np.full(4, 0.03)], 1)      # a thin chain joining the first two blobs
    pts = np.concatenate(blobs + [bridge, rng.uniform(-3, 8, size=(500, 2))]).astype(np.float32)
    pts = pts[rng.permutation(len(pts))]
    assert np.array_equal(pipeline.group_dbscan(pts, 0.15, 50, -1, 1), cluster_ref.group_dbscan_ref(pts, 0.15, 50, -1, 1))
