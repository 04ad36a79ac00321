"""GPU tests of the halo lists (tl_halo_build) and of the halo-cached submanifold convolution (csrc/tl_conv_halo.cu), the
kernel behind modes f16 / f16x2 for 3^3 submanifold layers.  The conv parity tests of test_gpu_ts.py / test_gpu_path.py
already run through it (levels built by sparse.build_levels carry the lists); here: the lists themselves against numpy,
the halo kernel against the gather kernel on the same inputs, a two-segment layer, a multi-tile level with a ragged last
tile, and the fallback when a tile's list exceeds the cap."""
import numpy as np
import pytest
import torch

from treelearn_b200 import sparse, synth, _lib

pytestmark = pytest.mark.gpu
SHAPE = [500, 500, 1000]


def _level(edge=8.0, seed=11):
    batch = synth.make_batch([synth.synth_forest(edge=edge, n_trees=4, seed=seed, ground_density=400.0)])
    dev = 'cuda'
    vf, vc, keys, v2p = sparse.voxelize(batch['coords'].to(dev), batch['input_feats'].to(dev), batch['batch_ids'].to(dev),
                                        batch['batch_size'], 0.1, False, True, 3)
    return sparse.build_levels(keys, vc, SHAPE, 1)[0]


def test_halo_lists_match_numpy():
    """The lists against numpy: the distinct neighbour rows of every tile, each at a position whose parity is bit 2 of its
    parity class; the rulebook as (position | class << 13) per LANE with the lane -> row permutation in row 27; and the
    purpose of it all -- within every 8-lane shared-memory phase the live entries of an offset carry different classes
    (= different bank groups) except where a tile holds more than 16 rows of one class."""
    lv = _level()
    h = lv.nbr.halo
    assert h.usable and h.umax <= h.cap
    nbr = lv.nbr.cpu().numpy()[:, :lv.n]
    keys = lv.keys.cpu().numpy()
    rows, cnt, lidx = h.rows.cpu().numpy(), h.cnt.cpu().numpy(), h.lidx.cpu().numpy().astype(np.int64) & 0xffff
    tiles = (lv.n + 127) // 128
    assert tiles > 3 and lv.n % 128 != 0            # several tiles, ragged last one
    umax = 0
    live_total = clash_total = natural_clash = 0
    for t in range(tiles):
        blk = nbr[:, t * 128:(t + 1) * 128]
        ncols = blk.shape[1]
        want = np.unique(blk[blk >= 0])
        ent = rows[t, :cnt[t]]
        used = ent[ent >= 0]
        ids, cls = used & ((1 << 28) - 1), used >> 28
        assert np.array_equal(np.sort(ids), want)                           # every distinct neighbour row exactly once
        assert np.array_equal(cls, keys[ids] & 7)                           # tagged with its parity class
        pos = np.nonzero(ent >= 0)[0] + 1
        assert np.array_equal(pos & 1, cls >> 2)                            # odd position <=> class bit 2
        umax = max(umax, int(cnt[t]))
        perm = lidx[t][27]
        assert np.array_equal(np.sort(perm), np.arange(128))                # lane -> tile row permutation
        got = lidx[t][:27]                                                  # [k][lane]
        p, c = got & 0x1FFF, got >> 13
        for k in range(27):
            r = perm                                                        # tile row of every lane
            inside = r < ncols
            v = np.where(inside, blk[k, np.minimum(r, ncols - 1)], -1)
            assert np.array_equal(p[k] == 0, v < 0)                         # 0 <=> absent neighbour (or a row past n_out)
            live = v >= 0
            assert np.array_equal(ent[p[k][live] - 1] & ((1 << 28) - 1), v[live])
            assert np.array_equal(c[k][live], keys[v[live]] & 7)
            for ph in range(16):                                            # bank-group clashes inside the 8-lane phases
                sel = live[8 * ph:8 * ph + 8]
                cc = c[k][8 * ph:8 * ph + 8][sel]
                live_total += len(cc)
                clash_total += len(cc) - len(set(cc.tolist()))
                nat = blk[k, 8 * ph:min(8 * ph + 8, ncols)]
                nat = keys[nat[nat >= 0]] & 7
                natural_clash += len(nat) - len(set(nat.tolist()))
    assert h.umax == umax
    print(f'same-class entries inside a phase: {clash_total} of {live_total} live entries with the lane permutation, '
          f'{natural_clash} in natural row order')
    assert clash_total < 0.12 * live_total and clash_total < 0.5 * natural_clash


@pytest.mark.parametrize('nsplit', [1, 2])
@pytest.mark.parametrize('ci,co,two_seg', [(32, 32, False), (64, 64, False), (32, 32, True), (96, 64, False), (128, 160, False)])
def test_halo_conv_equals_gather_conv(nsplit, ci, co, two_seg, monkeypatch):
    lv = _level()
    g = torch.Generator(device='cuda').manual_seed(ci + co + nsplit)
    mode = _lib.MODE_F16 if nsplit == 1 else _lib.MODE_F16X2
    prep = (lambda t: sparse.to_p(t).half().contiguous()) if nsplit == 1 else sparse.to_split

    def make_seg():
        x = torch.randn((lv.n, ci), device='cuda', generator=g)
        w = torch.randn((27, co, ci), device='cuda', generator=g) / (27 * ci) ** 0.5
        return prep(x), sparse.pack_weight(w, nsplit)
    srcs = [make_seg() for _ in range(2 if two_seg else 1)]
    res = torch.randn((lv.n, co), device='cuda', generator=g)
    s, t = torch.rand(co, device='cuda', generator=g) + 0.5, torch.randn(co, device='cuda', generator=g)

    def run():
        return sparse.conv([sparse.Seg(x, w, lv.nbr, lv.nbr_mask) for x, w in srcs], lv.n, co, mode, residual=res, raw=True,
                           act1=(s, t))
    lib = _lib.load()
    raw_h, act_h = run()                                  # halo lists attached to lv.nbr -> tl_conv_halo.cu
    h = lv.nbr.halo
    del lv.nbr.halo
    raw_g, act_g = run()                                  # no lists -> the gather kernel (tl_conv_grp.cu)
    lv.nbr.halo = h
    assert torch.allclose(raw_h, raw_g, atol=2e-4 if nsplit == 1 else 2e-5, rtol=1e-4), (raw_h - raw_g).abs().max()
    unpack = (lambda a: a.float()) if nsplit == 1 else sparse.from_split      # f16x2: compare hi + lo, not the halves
    assert torch.allclose(unpack(act_h), unpack(act_g), atol=2e-3 if nsplit == 1 else 1e-4, rtol=2e-3 if nsplit == 1 else 1e-4)


def test_halo_overflow_falls_back(monkeypatch):
    lv = _level()
    h = sparse.build_halo(lv, cap=64)                      # far below the ~250 distinct rows of a tile
    assert h.umax > 64 and not h.usable
    x = torch.randn((lv.n, 32), device='cuda').half()
    w = sparse.pack_weight(torch.randn((27, 32, 32), device='cuda') / 30, 1)
    out = sparse.conv([sparse.Seg(x, w, lv.nbr, lv.nbr_mask)], lv.n, 32, _lib.MODE_F16, raw=True)     # gather kernel
    lv2 = _level()
    out2 = sparse.conv([sparse.Seg(x, w, lv2.nbr, lv2.nbr_mask)], lv.n, 32, _lib.MODE_F16, raw=True)  # halo kernel
    assert torch.allclose(out, out2, atol=2e-4, rtol=1e-4)
