"""Host side of the hash-indexed sparse voxel tensor: point->voxel, level pyramid, rulebooks.

Replaces what `spconv.SparseConvTensor` + `indice_dict` carry in the reference
(tree_learn/model/tree_learn.py:88).  Rows of every level are kept
in Morton order (batch-major), so the 2^3 children of a coarse voxel are adjacent rows and
3^3 neighbours are close in memory; the reference leaves the row order implementation-defined.
torch only provides device memory and the stream.
"""
import ctypes as C
import os
from dataclasses import dataclass, field
from typing import List, Optional

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr, TILE_ROWS


def pad_rows(n):
    return (n + TILE_ROWS - 1) // TILE_ROWS * TILE_ROWS


def _workspace(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


@dataclass
class Level:
    n: int                       # active voxels
    shape: List[int]             # spatial shape (x,y,z) at this level
    keys: torch.Tensor           # [n] int64 view of the u64 Morton keys, ascending
    coords: torch.Tensor         # [n,4] int32 (b,x,y,z)
    nbr: Optional[torch.Tensor] = None        # [27, pad(n)] int32
    nbr_mask: Optional[torch.Tensor] = None   # [pad(n)/128] int32 bitmask
    down_index: Optional[torch.Tensor] = None  # [8, stride] rows of THIS level feeding the next (coarser) level
    down_mask: Optional[torch.Tensor] = None
    up_index: Optional[torch.Tensor] = None    # [8, stride] rows of the next level feeding THIS level
    up_mask: Optional[torch.Tensor] = None
    pair_stride: int = 0

    @property
    def nbr_stride(self):
        return pad_rows(self.n)


def voxelize(coords, input_feats, batch_ids, batch_size, voxel_size=0.1, use_coords=False, use_feats=True,
             max_num_points_per_voxel=3):
    """coords [N,3] f32, input_feats [N,F] f32, batch_ids [N] i64 ascending (all CUDA).
    Returns (voxel_feats [M,F+3] as [feat...,x,y,z], voxel_coords [M,4] i32, keys [M] i64, v2p [N] i64)."""
    lib = _lib.load()
    _lib.require_cuda(coords, input_feats, batch_ids)
    coords = coords.contiguous().float()
    n = coords.shape[0]
    f = 0 if input_feats is None else input_feats.shape[1]
    feats = None if f == 0 else input_feats.contiguous().float()
    bids = batch_ids.contiguous().long()
    dev = coords.device
    keys = torch.empty(n, dtype=torch.int64, device=dev)
    vcoords = torch.empty((n, 4), dtype=torch.int32, device=dev)
    vfeats = torch.empty((n, f + 3), dtype=torch.float32, device=dev)
    v2p = torch.empty(n, dtype=torch.int64, device=dev)
    m = C.c_int64(0)
    wsb = lib.tl_voxelize_workspace_bytes(n)
    ws = _workspace(wsb, dev)
    check(lib.tl_voxelize(ptr(coords), ptr(feats), f, ptr(bids), n, int(batch_size), float(voxel_size),
                          int(bool(use_coords)), int(bool(use_feats)), int(max_num_points_per_voxel),
                          ptr(keys), ptr(vcoords), ptr(vfeats), ptr(v2p), C.byref(m), ptr(ws), wsb, stream_ptr()))
    m = m.value
    return vfeats[:m], vcoords[:m], keys[:m], v2p


def build_levels(keys, coords, spatial_shape, num_levels, subm=True):
    """Level pyramid (strided k2/s2 maps) + 3^3 rulebook per level.  Raises ValueError('... reach zero!!! ...')
    like spconv when an axis of the U-Net collapses (tree_learn/util/pipeline.py:91-97)."""
    levels = [Level(n=int(keys.shape[0]), shape=[int(s) for s in spatial_shape], keys=keys, coords=coords)]
    _extend_levels(levels, num_levels)
    if subm:
        for lv in levels:
            build_subm_rulebook(lv)
        read_halo_sizes(levels)
    return levels


def read_halo_sizes(levels):
    """One host read for the largest halo list of EVERY level (instead of one synchronisation per level in the middle of the
    conv stream, where the host would wait for all queued convolutions before it can enqueue the next level's)."""
    halos = [lv.nbr.halo for lv in levels if lv.nbr is not None and getattr(lv.nbr, 'halo', None) is not None
             and lv.nbr.halo._umax is None]
    if halos:
        for h, v in zip(halos, torch.cat([h.max_cnt for h in halos]).tolist()):
            h._umax = int(v)


def _extend_levels(levels, num_levels):
    """Append the coarser levels (strided maps of levels[-1] ... ) until there are `num_levels`."""
    lib = _lib.load()
    dev = levels[0].keys.device
    for l in range(len(levels) - 1, num_levels - 1):
        fine = levels[-1]
        n = fine.n
        stride = pad_rows(n)
        ckeys = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
        ccoords = torch.empty((max(n, 1), 4), dtype=torch.int32, device=dev)
        fine.down_index = torch.empty((8, stride), dtype=torch.int32, device=dev)
        fine.up_index = torch.empty((8, stride), dtype=torch.int32, device=dev)
        fine.down_mask = torch.empty(stride // TILE_ROWS, dtype=torch.int32, device=dev)
        fine.up_mask = torch.empty(stride // TILE_ROWS, dtype=torch.int32, device=dev)
        fine.pair_stride = stride
        fshape = (C.c_int32 * 3)(*fine.shape)
        cshape = (C.c_int32 * 3)()
        nc = C.c_int64(0)
        if n == 0:
            cs = [(s - 2) // 2 + 1 for s in fine.shape]
            if min(cs) <= 0:
                raise ValueError(f'your out spatial shape {cs} reach zero!!! input shape: {fine.shape}')
            levels.append(Level(n=0, shape=cs, keys=ckeys[:0], coords=ccoords[:0]))
            continue
        wsb = lib.tl_level_workspace_bytes(n)
        ws = _workspace(wsb, dev)
        check(lib.tl_build_level(ptr(fine.keys), n, fshape, ptr(ckeys), ptr(ccoords), ptr(fine.down_index),
                                 ptr(fine.down_mask), ptr(fine.up_index), ptr(fine.up_mask), cshape, C.byref(nc),
                                 ptr(ws), wsb, stream_ptr()))
        levels.append(Level(n=nc.value, shape=list(cshape), keys=ckeys[:nc.value], coords=ccoords[:nc.value]))
    return levels


class LazyLevels:
    """Level pyramid whose coarser levels are built late and on a side stream.

    Level 0 and its 3^3 rulebook are built at construction on the current stream.  The strided maps and the rulebooks of
    levels >= 1 are built on first access of any level >= 1 -- by then the caller has already enqueued the level-0
    convolutions on the main stream, so the ~1.2 ms of hash / scan / probe kernels (and their host synchronisations) run
    concurrently with them on the side stream; the main stream then waits on one event.  (The conv kernel keeps one CTA
    per SM with ~190 KB of shared memory; the small geometry kernels fit beside it.)"""
    _side = {}

    def __init__(self, keys, coords, spatial_shape, num_levels):
        self.num_levels = num_levels
        self.levels = [Level(n=int(keys.shape[0]), shape=[int(s) for s in spatial_shape], keys=keys, coords=coords)]
        build_subm_rulebook(self.levels[0])
        self._done = num_levels <= 1

    def _finish(self):
        if self._done:
            return
        self._done = True
        main = torch.cuda.current_stream()
        dev = self.levels[0].keys.device
        side = LazyLevels._side.get(dev.index)
        if side is None:
            side = LazyLevels._side[dev.index] = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(side):
            _extend_levels(self.levels, self.num_levels)
            for lv in self.levels[1:]:
                build_subm_rulebook(lv)
            for lv in self.levels[1:]:      # one host read per level, on the side stream: the main stream keeps running
                h = getattr(lv.nbr, 'halo', None)
                if h is not None:
                    h.umax
        main.wait_stream(side)
        for lv in self.levels:        # tensors allocated under the side stream are consumed by main-stream kernels
            for t in (lv.keys, lv.coords, lv.nbr, lv.nbr_mask, lv.down_index, lv.down_mask, lv.up_index, lv.up_mask):
                if t is not None:
                    t.record_stream(main)
            h = getattr(lv.nbr, 'halo', None) if lv.nbr is not None else None
            if h is not None:
                for t in (h.rows, h.cnt, h.lidx):
                    t.record_stream(main)

    def __len__(self):
        return self.num_levels

    def __getitem__(self, l):
        if l != 0:
            self._finish()
        return self.levels[l]

    def __iter__(self):
        self._finish()
        return iter(self.levels)


def build_subm_rulebook(lv):
    lib = _lib.load()
    dev = lv.keys.device
    stride = pad_rows(lv.n)
    lv.nbr = torch.empty((27, max(stride, TILE_ROWS)), dtype=torch.int32, device=dev)
    lv.nbr_mask = torch.empty(max(stride // TILE_ROWS, 1), dtype=torch.int32, device=dev)
    if lv.n == 0:
        return lv
    wsb = lib.tl_rulebook_workspace_bytes(lv.n)
    ws = _workspace(wsb, dev)
    shape = (C.c_int32 * 3)(*lv.shape)
    check(lib.tl_subm_rulebook(ptr(lv.keys), lv.n, shape, ptr(lv.nbr), ptr(lv.nbr_mask), ptr(ws), wsb, stream_ptr()))
    if USE_HALO:
        build_halo(lv)
    return lv


class Halo:
    """Per 128-row tile of a level: the distinct neighbour rows (`rows`: row id | parity class << 28 per list position), the
    3^3 rulebook as 16-bit entries into that list per lane and the lane -> tile row permutation (`lidx` [tiles, 28, 128])
    (tl_halo_build, csrc/tl_conv_halo.cu).  Hangs off the level's `nbr` tensor so that sparse.conv finds it from the
    segments' index tensors.  `umax` (largest list of the level) costs one host read of an int32."""

    def __init__(self, rows, cnt, lidx, max_cnt, cap):
        self.rows, self.cnt, self.lidx, self.max_cnt, self.cap, self._umax = rows, cnt, lidx, max_cnt, cap, None

    @property
    def umax(self):
        if self._umax is None:
            self._umax = int(self.max_cnt.item())
        return self._umax

    @property
    def usable(self):
        return 0 < self.umax <= self.cap


def build_halo(lv, cap=None):
    lib = _lib.load()
    dev = lv.keys.device
    cap = cap or HALO_CAP
    tiles = pad_rows(lv.n) // TILE_ROWS
    h = Halo(torch.empty((tiles, cap), dtype=torch.int32, device=dev), torch.empty(tiles, dtype=torch.int32, device=dev),
             torch.empty((tiles, 28, TILE_ROWS), dtype=torch.int16, device=dev), torch.zeros(1, dtype=torch.int32, device=dev), cap)
    check(lib.tl_halo_build(ptr(lv.nbr), ptr(lv.keys), lv.n, lv.nbr.stride(0), cap, ptr(h.rows), ptr(h.cnt), ptr(h.lidx),
                            ptr(h.max_cnt), stream_ptr()))
    lv.nbr.halo = h
    return h


# ---- segmented gather-GEMM convolution -----------------------------------------------------------
@dataclass
class Seg:
    src: torch.Tensor                      # [rows, c_in] f32 contiguous
    weight: torch.Tensor                   # packed for the active mode, [n_off, ...]
    index: Optional[torch.Tensor] = None   # [n_off, stride] i32 or None (identity)
    mask: Optional[torch.Tensor] = None


def round_tf32(w):
    """Round-to-nearest-even to TF32 (10 explicit mantissa bits) so the tensor core's operand truncation is exact."""
    i = w.contiguous().view(torch.int32)
    return ((i + 0x0FFF + ((i >> 13) & 1)) & ~0x1FFF).view(torch.float32)


def pack_weight_tc(w, half, bk=None):
    """w [n_off, C_out, C_in] -> the tcgen05 path's B-operand layout [n_off, C_in/bk, C_out, bk]: one contiguous
    [C_out x bk-channel] slab per (offset, k-block) whose rows already carry the UMMA shared-memory swizzle (16 B chunk
    c of row n sits at c ^ (n & 7) for 128 B rows, at c ^ ((n >> 1) & 3) for 64 B rows), so the kernel lands a slab in
    its B stage with a single TMA bulk copy.  fp16 when `half`, else fp32 rounded to TF32.  bk = channels per chunk:
    32, or 64 for fp16 weights whose C_in is a multiple of 64 (the rule csrc/tl_conv_tc.cu::conv_fwd_tc applies)."""
    k, co, ci = w.shape
    assert ci % 32 == 0 and co % 32 == 0, (ci, co)
    if bk is None:
        bk = 64 if (half and ci % 64 == 0 and os.environ.get('TL_TC_BK64', '1') != '0') else 32
    t = w.detach().half() if half else round_tf32(w.detach().float())
    row_bytes = bk * (2 if half else 4)
    ch = row_bytes // 16                       # 16 B chunks per row
    t = t.reshape(k, co, ci // bk, ch, bk // ch).permute(0, 2, 1, 3, 4)      # [k, kb, co, chunk, elems]
    n = torch.arange(co, device=w.device)
    x = (n & 7) if row_bytes == 128 else ((n >> 1) & 3)
    src_chunk = torch.arange(ch, device=w.device)[None, :] ^ x[:, None]     # destination chunk c holds source chunk c ^ x
    t = torch.gather(t, 3, src_chunk[None, None, :, :, None].expand(k, ci // bk, co, ch, bk // ch))
    return t.reshape(k, ci // bk, co, bk).contiguous()


# TL_TS: 1 (default) group kernel csrc/tl_conv_grp.cu; 2 tensor-memory-A kernel csrc/tl_conv_ts.cu; 0 round-1 kernel (natural layout)
TS_KIND = int(os.environ.get('TL_TS', '1'))
USE_TS = TS_KIND != 0
# halo-cached submanifold conv (csrc/tl_conv_halo.cu) behind the group kernel's modes; TL_HALO=0 turns it off
USE_HALO = TS_KIND == 1 and os.environ.get('TL_HALO', '1') != '0'
HALO_CAP = int(os.environ.get('TL_HALO_CAP', '512'))


# ---- "P-layout" of the tensor-memory-A kernel (csrc/tl_conv_ts.cu) -----------------------------------------------------
# Within every 32-channel block, position m = 8q + 2g + e of a row holds logical channel 8g + 2q + e (q, g in 0..3,
# e in 0..1): PI[m] = logical channel stored at position m.  Used by the fp32 residual stream and the activated operand
# tensors of modes f16 / f16x2; scale / shift vectors and the model's inputs / outputs stay in logical order.
PI = [8 * ((m % 8) // 2) + 2 * (m // 8) + (m % 2) for m in range(32)]
PI_INV = [PI.index(c) for c in range(32)]


def to_p(x):
    """[rows, C] logical channel order -> P-layout (C % 32 == 0)."""
    r, c = x.shape
    return x.reshape(r, c // 32, 32)[:, :, torch.tensor(PI, device=x.device)].reshape(r, c).contiguous()


def from_p(x):
    r, c = x.shape
    return x.reshape(r, c // 32, 32)[:, :, torch.tensor(PI_INV, device=x.device)].reshape(r, c).contiguous()


def pack_weight_ts(w, nsplit=1):
    """w [n_off, C_out, C_in] (logical channels) -> the B-operand image of the tensor-memory-A kernel:
    [n_off, C_in/32, nsplit, C_out, 32] fp16, one contiguous slab per (offset, 32-channel block).
    * K order inside a block: K step kk, position 4q + j is what lane q of a quad delivers from bytes [8kk, 8kk + 8) of its
      16 B piece of a P-layout row, i.e. memory position 8q + 4kk + j = logical channel PI[8q + 4kk + j];
    * every 64 B row carries the UMMA SWIZZLE_64B image (16 B chunk c of row n at c ^ ((n >> 1) & 3));
    * nsplit = 2 (mode f16x2): a (hi, lo) pair of slabs with hi = fp16(w), lo = fp16(w - hi)."""
    k, co, ci = w.shape
    assert ci % 32 == 0 and co % 32 == 0, (ci, co)
    w = w.detach().float()
    hi = w.half()
    parts = [hi] if nsplit == 1 else [hi, (w - hi.float()).half()]
    dev = w.device
    chan = torch.tensor([PI[8 * ((p % 16) // 4) + 4 * (p // 16) + (p % 4)] for p in range(32)], device=dev)
    n = torch.arange(co, device=dev)
    src_chunk = torch.arange(4, device=dev)[None, :] ^ ((n >> 1) & 3)[:, None]   # destination chunk c holds source chunk c ^ x
    out = []
    for t in parts:
        t = t.reshape(k, co, ci // 32, 32)[..., chan]              # K order of the kernel
        t = t.permute(0, 2, 1, 3).reshape(k, ci // 32, co, 4, 8)   # [k, kb, co, chunk, 8 halves]
        t = torch.gather(t, 3, src_chunk[None, None, :, :, None].expand(k, ci // 32, co, 4, 8))
        out.append(t.reshape(k, ci // 32, co, 32))
    return torch.stack(out, 2).contiguous()                        # [k, kb, nsplit, co, 32]


def pack_weight_grp(w, nsplit=1):
    """w [n_off, C_out, C_in] (logical channels) -> the B-operand image of the group kernel (csrc/tl_conv_grp.cu):
    [n_off, C_in/32, nsplit, C_out, 32] fp16; K position m of a 32-channel block = logical channel PI[m] (the memory order of
    a P-layout row, which cp.async copies verbatim), rows carry the SWIZZLE_64B image; nsplit = 2: (hi, lo) slab pairs."""
    k, co, ci = w.shape
    assert ci % 32 == 0 and co % 32 == 0, (ci, co)
    w = w.detach().float()
    hi = w.half()
    parts = [hi] if nsplit == 1 else [hi, (w - hi.float()).half()]
    dev = w.device
    chan = torch.tensor(PI, device=dev)
    n = torch.arange(co, device=dev)
    src_chunk = torch.arange(4, device=dev)[None, :] ^ ((n >> 1) & 3)[:, None]
    out = []
    for t in parts:
        t = t.reshape(k, co, ci // 32, 32)[..., chan]
        t = t.permute(0, 2, 1, 3).reshape(k, ci // 32, co, 4, 8)
        t = torch.gather(t, 3, src_chunk[None, None, :, :, None].expand(k, ci // 32, co, 4, 8))
        out.append(t.reshape(k, ci // 32, co, 32))
    return torch.stack(out, 2).contiguous()


def pack_weight(w, nsplit=1):
    """Weights for the active f16 / f16x2 kernel (TL_TS)."""
    return pack_weight_grp(w, nsplit) if TS_KIND == 1 else pack_weight_ts(w, nsplit)


def permute_p_weight(w):
    """w [n_off, C_out, C_in] logical -> both channel axes in P-layout order, for kernels with natural addressing (fp32 SIMT,
    round 1's TF32 kernel) that read and write P-layout tensors (the 1x1 projection of the fp32 residual stream)."""
    k, co, ci = w.shape
    pi = torch.tensor(PI, device=w.device)
    w = w.reshape(k, co // 32, 32, ci // 32, 32)[:, :, pi][:, :, :, :, pi]
    return w.reshape(k, co, ci).contiguous()


def to_split(x):
    """fp32 [rows, C] logical -> the f16x2 operand format: P-layout, per 32-channel block 64 B of hi halves then 64 B of lo
    halves ([rows, C/32, 2, 32] fp16, returned as [rows, 2C])."""
    x = to_p(x)
    r, c = x.shape
    hi = x.half()
    lo = (x - hi.float()).half()
    return torch.stack([hi.reshape(r, c // 32, 32), lo.reshape(r, c // 32, 32)], 2).reshape(r, 2 * c).contiguous()


def from_split(x):
    r, c2 = x.shape
    v = x.reshape(r, c2 // 64, 2, 32).float()
    return from_p((v[:, :, 0] + v[:, :, 1]).reshape(r, c2 // 2))


def split_to_half(x):
    """f16x2 operand format [rows, 2C] (per 32-channel block: 32 hi halves, 32 lo halves) -> fp16 [rows, C] = fp16(hi + lo),
    same P-layout channel order."""
    r, c2 = x.shape
    v = x.reshape(r, c2 // 64, 2, 32)
    return (v[:, :, 0].float() + v[:, :, 1].float()).half().reshape(r, c2 // 2).contiguous()


def half_to_split(x):
    """fp16 [rows, C] -> f16x2 operand format [rows, 2C] with zero lo terms."""
    r, c = x.shape
    v = x.reshape(r, c // 32, 1, 32)
    return torch.cat([v, torch.zeros_like(v)], 2).reshape(r, 2 * c).contiguous()


SPLITK_MAX_ROWS = 2 * 148 * TILE_ROWS   # below two waves of 128-row tiles the library may split K over CTAs
PROFILE = None   # bench.py sets this to a list: every conv launch appends (start_evt, end_evt, alg_bytes, flops)


def conv(segs, n_out, c_out, mode, residual=None, raw=False, act1=None, act2=None):
    """acc = sum_seg sum_k W[k] . src[index[k]] (+ residual); returns (raw?, act1?, act2?) tensors that were
    asked for, in that order.  act = (scale, shift) per channel => relu(scale*v+shift)  (BN eval + ReLU)."""
    lib = _lib.load()
    dev = segs[0].src.device
    d = _lib.ConvDesc()
    d.n_out, d.c_out, d.n_seg = int(n_out), int(c_out), len(segs)
    half_modes = (_lib.MODE_F16, _lib.MODE_F16X2)
    for i, s in enumerate(segs):
        g = d.seg[i]
        g.src, g.src_stride, g.c_in = ptr(s.src), s.src.stride(0), s.src.shape[1]
        if mode in half_modes and s.src.dtype == torch.float32:
            d.src_fp32_mask |= 1 << i          # raw residual-stream rows, converted to the operand format in registers
        elif mode == _lib.MODE_F16X2:          # [rows, C/32, 2, 32] fp16 stored as [rows, 2C]
            g.src_stride, g.c_in = s.src.stride(0) // 2, s.src.shape[1] // 2
        g.weight, g.n_off = ptr(s.weight), s.weight.shape[0]
        if s.index is not None:
            g.index, g.index_stride, g.tile_mask = ptr(s.index), s.index.stride(0), ptr(s.mask)
    halo = getattr(segs[0].index, 'halo', None) if (mode in half_modes and not d.src_fp32_mask) else None
    if halo is not None and all(sg.index is segs[0].index for sg in segs) and halo.usable:
        d.halo_rows, d.halo_cnt, d.halo_lidx = ptr(halo.rows), ptr(halo.cnt), ptr(halo.lidx)
        d.halo_cap, d.halo_umax = halo.cap, halo.umax
    outs = []
    d.residual = ptr(residual)
    act_dtype = torch.float16 if mode in half_modes else torch.float32   # operand format of the consumers
    act_cols = 2 * c_out if mode == _lib.MODE_F16X2 else c_out            # f16x2: (hi, lo) halves per channel
    if raw:
        o = torch.empty((n_out, c_out), dtype=torch.float32, device=dev)
        d.out_raw = ptr(o)
        outs.append(o)
    if act1 is not None:
        o = torch.empty((n_out, act_cols), dtype=act_dtype, device=dev)
        d.out_act1, d.scale1, d.shift1 = ptr(o), ptr(act1[0]), ptr(act1[1])
        outs.append(o)
    if act2 is not None:
        o = torch.empty((n_out, act_cols), dtype=act_dtype, device=dev)
        d.out_act2, d.scale2, d.shift2 = ptr(o), ptr(act2[0]), ptr(act2[1])
        outs.append(o)
    if mode != _lib.MODE_FP32 and 0 < n_out < SPLITK_MAX_ROWS:
        ws = torch.empty((n_out, c_out), dtype=torch.float32, device=dev)   # deep levels: few tiles -> split-K scratch
        d.splitk_ws = ptr(ws)
    if PROFILE is not None and n_out > 0:
        # algorithmic traffic: every input row once + one output + index tables + weights
        # ONE output per conv in the element size of the stream it feeds (extra activated copies, the residual
        # read and BN/ReLU count as fused = 0 bytes); f16x2 activations are 4 B per channel (hi + lo)
        byts = n_out * c_out * (2 if (mode == _lib.MODE_F16 and not raw) else 4)
        flops = 0
        for s in segs:
            byts += s.src.shape[0] * s.src.shape[1] * s.src.element_size() + s.weight.numel() * s.weight.element_size()
            if s.index is not None:
                byts += s.weight.shape[0] * n_out * 4
            flops += 2 * s.weight.shape[0] * n_out * s.src.shape[1] * c_out    # dense upper bound (all offsets present)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(lib.tl_conv_fwd(C.byref(d), int(mode), stream_ptr()))
        e1.record()
        PROFILE.append((e0, e1, byts, flops, c_out))
    else:
        check(lib.tl_conv_fwd(C.byref(d), int(mode), stream_ptr()))
    return outs[0] if len(outs) == 1 else tuple(outs)


def heads(voxel_feats, v2p, packed, split=False):
    """voxel->point gather + both 2-layer heads (tree_learn.py:97-103).  packed: dict of folded weights.
    voxel_feats: fp32 [M,C], fp16 [M,C], or (split=True) the f16x2 operand format [M,2C]."""
    lib = _lib.load()
    n = int(v2p.shape[0])
    c = int(voxel_feats.shape[1]) // (2 if split else 1)
    dev = voxel_feats.device
    feats = torch.empty((n, c), dtype=torch.float32, device=dev)
    logits = torch.empty((n, 2), dtype=torch.float32, device=dev)
    offs = torch.empty((n, 3), dtype=torch.float32, device=dev)
    fmt = 0 if voxel_feats.dtype == torch.float32 else (2 if split else (3 if USE_TS else 1))
    check(lib.tl_heads_fwd(ptr(voxel_feats), fmt, ptr(v2p), n, c, ptr(packed['sem_w1']), ptr(packed['sem_b1']),
                           ptr(packed['sem_w2']), ptr(packed['sem_b2']), ptr(packed['off_w1']), ptr(packed['off_b1']),
                           ptr(packed['off_w2']), ptr(packed['off_b2']), ptr(feats), ptr(logits), ptr(offs),
                           stream_ptr()))
    return feats, logits, offs
