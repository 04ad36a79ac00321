"""GPU parity tests of the f16 / f16x2 convolution kernels (csrc/tl_conv_grp.cu; csrc/tl_conv_ts.cu for raw fp32 sources)
through the C ABI:
mode f16 (fp16 operands) and mode f16x2 (two-term fp16 split = fp32-equivalent products) against the oracle's
rulebook convolutions (oracle/model_ref.py, which restates spconv's SubMConv3d / SparseConv3d / SparseInverseConv3d:
reference tree_learn/model/blocks.py:57-70,104-123)."""
import numpy as np
import pytest
import torch

from oracle import model_ref, spconv_ref as sp
from treelearn_b200 import TreeLearn, sparse, synth, _lib

pytestmark = pytest.mark.gpu
SHAPE = [500, 500, 1000]
F16_EXACT = dict(atol=3e-4, rtol=2e-4)      # fp16-exact operands: only the fp32 accumulation order differs
X2_TOL = dict(atol=6e-5, rtol=6e-5)         # f16x2: 22-bit operands, fp32 accumulation in the tensor core (up to 6 000 terms) -- vs float64


def _geom(seed=9, levels=1):
    batch = synth.make_batch([synth.synth_forest(edge=5.0, n_trees=2, seed=seed, ground_density=200.0)])
    dev = 'cuda'
    vf, vc, keys, v2p = sparse.voxelize(batch['coords'].to(dev), batch['input_feats'].to(dev), batch['batch_ids'].to(dev),
                                        batch['batch_size'], 0.1, False, True, 3)
    return vc, sparse.build_levels(keys, vc, SHAPE, levels)


def _wk(w, co, k, ci):
    return w.reshape(co, k, ci).permute(1, 0, 2).cuda().float()


@pytest.mark.parametrize('nsplit', [1, 2])
@pytest.mark.parametrize('ci,co', [(32, 32), (64, 32), (32, 64), (64, 64), (96, 96), (128, 128), (160, 160), (224, 224)])
def test_ts_subm_conv_parity(nsplit, ci, co):
    vc, (lv,) = _geom()
    g = torch.Generator().manual_seed(ci * 1000 + co + nsplit)
    x = torch.randn((lv.n, ci), generator=g)
    w = torch.randn((co, 3, 3, 3, ci), generator=g) / (27 * ci) ** 0.5
    if nsplit == 1:
        x, w = x.half().float(), w.half().float()
    res = torch.randn((lv.n, co), generator=g)
    s, t = torch.rand(co, generator=g) + 0.5, torch.randn(co, generator=g)
    nbr = sp.subm_neighbour_table(vc.cpu().numpy(), SHAPE)
    ref = (model_ref._subm(x.double(), nbr, w.double()) + res.double())
    mode = _lib.MODE_F16 if nsplit == 1 else _lib.MODE_F16X2
    xs = sparse.to_p(x.cuda()).half() if nsplit == 1 else sparse.to_split(x.cuda())
    raw, act, act2 = sparse.conv([sparse.Seg(xs, sparse.pack_weight(_wk(w, co, 27, ci), nsplit), lv.nbr, lv.nbr_mask)],
                                 lv.n, co, mode, residual=sparse.to_p(res.cuda()), raw=True, act1=(s.cuda(), t.cuda()),
                                 act2=(t.cuda().abs() + 0.1, s.cuda()))
    tol = F16_EXACT if nsplit == 1 else X2_TOL
    assert raw.dtype == torch.float32 and act.dtype == torch.float16
    raw = sparse.from_p(raw)        # the kernel's tensors are in P-layout (treelearn_b200/sparse.py)
    assert torch.allclose(raw.cpu().double(), ref, **tol), (raw.cpu().double() - ref).abs().max()
    a1, a2 = torch.relu(ref * s + t), torch.relu(ref * (t.abs() + 0.1) + s)
    if nsplit == 1:
        assert torch.allclose(sparse.from_p(act.float()).cpu().double(), a1, atol=3e-3, rtol=2e-3)      # + fp16 rounding of the store
        assert torch.allclose(sparse.from_p(act2.float()).cpu().double(), a2, atol=3e-3, rtol=2e-3)
    else:
        assert torch.allclose(sparse.from_split(act).cpu().double(), a1, atol=1.2e-4, rtol=6e-5)    # scale up to 1.5 on the raw error
        assert torch.allclose(sparse.from_split(act2).cpu().double(), a2, atol=4e-4, rtol=6e-5)   # scale up to ~5 amplifies the raw error


@pytest.mark.parametrize('nsplit', [1, 2])
def test_ts_strided_inverse_and_fp32_identity_segments(nsplit):
    vc, (lv, nx) = _geom(seed=10, levels=2)
    g = torch.Generator().manual_seed(3 + nsplit)
    rnd = lambda *shape: torch.randn(shape, generator=g)   # noqa: E731
    q = (lambda a: a.half().float()) if nsplit == 1 else (lambda a: a)
    mode = _lib.MODE_F16 if nsplit == 1 else _lib.MODE_F16X2
    fmt = (lambda a: sparse.to_p(a.cuda()).half()) if nsplit == 1 else (lambda a: sparse.to_split(a.cuda()))
    tol = F16_EXACT if nsplit == 1 else X2_TOL
    x, wd, wu = q(rnd(lv.n, 32)), q(rnd(64, 2, 2, 2, 32) / 16), q(rnd(32, 2, 2, 2, 64) / 16)
    out_idx, out_shape, in_row, kappa, out_row = sp.strided_pairs(vc.cpu().numpy(), SHAPE)
    where = {tuple(r): i for i, r in enumerate(nx.coords.cpu().numpy().tolist())}
    out_row = np.array([where[tuple(r)] for r in out_idx.tolist()])[out_row]
    ref_d = model_ref._pairs_conv(x.double(), wd.double(), in_row, kappa, out_row, nx.n)
    d = sparse.conv([sparse.Seg(fmt(x), sparse.pack_weight(_wk(wd, 64, 8, 32), nsplit), lv.down_index, lv.down_mask)],
                    nx.n, 64, mode, raw=True)
    assert torch.allclose(sparse.from_p(d).cpu().double(), ref_d, **tol)
    dq = q(ref_d.float())
    ref_u = model_ref._pairs_conv(dq.double(), wu.double(), out_row, kappa, in_row, lv.n)
    u = sparse.conv([sparse.Seg(fmt(dq), sparse.pack_weight(_wk(wu, 32, 8, 64), nsplit), lv.up_index, lv.up_mask)],
                    lv.n, 32, mode, raw=True)
    assert torch.allclose(sparse.from_p(u).cpu().double(), ref_u, **tol)
    # blocks_tail.block0 second conv: 3^3 conv over the activated tensor + the 1x1 projection of the RAW fp32 residual
    # stream (two identity segments whose fp32 rows are converted to the operand format in registers)
    h, z, e = q(rnd(lv.n, 32)), q(rnd(lv.n, 32)), q(rnd(lv.n, 32))
    w3, wz, we = q(rnd(32, 3, 3, 3, 32) / 32), q(rnd(32, 1, 1, 1, 32) / 8), q(rnd(32, 1, 1, 1, 32) / 8)
    nbr = sp.subm_neighbour_table(vc.cpu().numpy(), SHAPE)
    ref = (model_ref._subm(h.double(), nbr, w3.double()) + z.double() @ wz.reshape(32, 32).T.double()
           + e.double() @ we.reshape(32, 32).T.double())
    out = sparse.conv([sparse.Seg(fmt(h), sparse.pack_weight_ts(_wk(w3, 32, 27, 32), nsplit), lv.nbr, lv.nbr_mask),
                       sparse.Seg(sparse.to_p(z.cuda()), sparse.pack_weight_ts(_wk(wz, 32, 1, 32), nsplit)),
                       sparse.Seg(sparse.to_p(e.cuda()), sparse.pack_weight_ts(_wk(we, 32, 1, 32), nsplit))], lv.n, 32, mode, raw=True)
    assert torch.allclose(sparse.from_p(out).cpu().double(), ref, **tol)
    # two 3^3 segments sharing one rulebook (the skip concat of blocks_tail.block0's first conv), 64 -> 32
    h2, w3b = q(rnd(lv.n, 32)), q(rnd(32, 3, 3, 3, 32) / 32)
    ref2 = model_ref._subm(h.double(), nbr, w3.double()) + model_ref._subm(h2.double(), nbr, w3b.double())
    out2 = sparse.conv([sparse.Seg(fmt(h), sparse.pack_weight(_wk(w3, 32, 27, 32), nsplit), lv.nbr, lv.nbr_mask),
                        sparse.Seg(fmt(h2), sparse.pack_weight(_wk(w3b, 32, 27, 32), nsplit), lv.nbr, lv.nbr_mask)],
                       lv.n, 32, mode, raw=True)
    assert torch.allclose(sparse.from_p(out2).cpu().double(), ref2, **tol)


def test_f16x2_default_model_close_to_fp32_oracle():
    """mode f16x2 carries ~22 mantissa bits per operand: the whole 7-level model agrees with the fp32 oracle to fp32
    accumulation noise (the reference's inference arithmetic is fp32: configs/pipeline/pipeline.yaml:12 is never read)."""
    batch = synth.make_batch([synth.workload('tiny')])
    sd = model_ref.make_state_dict(channels=32, num_blocks=7, seed=0)
    net = TreeLearn(use_feats=False, use_coords=False, spatial_shape=SHAPE, mode='f16x2')
    net.load_state_dict(sd)
    net = net.cuda().eval()
    with torch.no_grad():
        mine = net(batch, return_loss=False)
        ref = model_ref.forward_ref(sd, batch, spatial_shape=SHAPE)
    eo = (mine['offset_predictions'].cpu() - ref['offset_predictions']).abs().max().item()
    el = (mine['semantic_prediction_logits'].cpu() - ref['semantic_prediction_logits']).abs().max().item()
    ef = (mine['backbone_feats'].cpu() - ref['backbone_feats']).abs().max().item()
    print('f16x2 max err: offsets', eo, 'logits', el, 'feats', ef)
    assert eo < 2e-5 and el < 5e-5 and ef < 1e-4
