"""Training path (SURVEY.md §8 a12 train mode, a16): autograd Functions over the C-ABI kernels.

The reference trains by torch autograd through spconv's conv modules and `nn.BatchNorm1d` applied to
`SparseConvTensor.features` (tree_learn/model/blocks.py:55-79, tools/training/train.py:32-44).  Here:

  * `sparse_conv`  forward  = tl_conv_fwd (raw output, no fused epilogue: BatchNorm needs batch statistics first)
                   dgrad    = tl_conv_fwd on the transposed rulebook with transposed weights
                              (3^3 table: same table, offsets mirrored k -> K-1-k; strided maps: down <-> up)
                   wgrad    = tl_conv_wgrad
  * `bn_relu`      forward  = tl_bn_stats + tl_bn_finalize (running-stat update, torch semantics) + tl_bn_relu_apply
                   backward = tl_bn_relu_bwd
  Residual adds, the skip concat, the voxel->point gather and the two tiny MLP heads stay torch ops (autograd
  handles them); so does the loss.
"""
import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib, sparse
from ._lib import check, ptr, stream_ptr
from .sparse import Seg


# TL_WGRAD_TC=0 falls back to the round-1 mma.sync + atomicAdd weight-gradient kernel (A/B measurements only)
USE_WGRAD_TC = os.environ.get('TL_WGRAD_TC', '1') != '0'


@dataclass
class ConvGeom:
    """One direction of a rulebook: rows of `n_in` feed rows of `n_out`; `*_t` is the transposed map."""
    n_in: int
    n_out: int
    index: Optional[torch.Tensor]      # [n_off, stride] over n_out rows (None = identity / 1x1)
    mask: Optional[torch.Tensor]
    index_t: Optional[torch.Tensor]    # [n_off, stride] over n_in rows
    mask_t: Optional[torch.Tensor]
    mirror: bool                       # True: transposed map is the same table with offsets mirrored (3^3 subm)


def subm_geom(lv):
    return ConvGeom(lv.n, lv.n, lv.nbr, lv.nbr_mask, lv.nbr, lv.nbr_mask, True)


def down_geom(fine, coarse):
    return ConvGeom(fine.n, coarse.n, fine.down_index, fine.down_mask, fine.up_index, fine.up_mask, False)


def up_geom(fine, coarse):
    return ConvGeom(coarse.n, fine.n, fine.up_index, fine.up_mask, fine.down_index, fine.down_mask, False)


def identity_geom(n):
    return ConvGeom(n, n, None, None, None, None, False)


def _tc_ok(c_in, c_out):
    return c_in % 32 == 0 and c_in <= 256 and c_out % 32 == 0 and c_out <= 256


def _round_tf32(w):
    i = w.contiguous().view(torch.int32)
    return ((i + 0x0FFF + ((i >> 13) & 1)) & ~0x1FFF).view(torch.float32)


def pack_param(weight, transpose, mirror, mode):
    """Conv parameter [C_out, k, k, k, C_in] -> (packed weights, kernel mode) for tl_conv_fwd.
    transpose=False: the forward conv; transpose=True: the data-gradient conv (C_in/C_out swapped, offsets mirrored when
    `mirror`).  tcgen05-eligible shapes are packed by one CUDA kernel (tl_pack_weight_tc, TF32-rounded fp32); the rest take
    the fp32 SIMT layout [K, C_in', C_out']."""
    co, ci = weight.shape[0], weight.shape[-1]
    k = w_noff(weight)
    co_p, ci_p = (ci, co) if transpose else (co, ci)
    w = weight.detach().float().contiguous()
    if mode != _lib.MODE_FP32 and _tc_ok(ci_p, co_p):
        out = torch.empty((k, ci_p // 32, co_p, 32), dtype=torch.float32, device=w.device)
        check(_lib.load().tl_pack_weight_tc(ptr(w), co, k, ci, int(transpose), int(mirror), 0, 32, ptr(out), stream_ptr()))
        return out, _lib.MODE_TF32
    w3 = w.reshape(co, k, ci)
    if not transpose:
        return w3.permute(1, 2, 0).contiguous(), _lib.MODE_FP32        # [K, C_in, C_out]
    wt = w3.permute(1, 0, 2)                                           # [K, C_in' = C_out, C_out' = C_in]
    if mirror:
        wt = wt.flip(0)
    return wt.contiguous(), _lib.MODE_FP32


class _SparseConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, weight, geom, mode):
        src = src.contiguous().float()
        co = weight.shape[0]
        wp, m = pack_param(weight, False, False, mode)
        out = sparse.conv([Seg(src, wp, geom.index, geom.mask)], geom.n_out, co, m, raw=True)
        ctx.save_for_backward(src, weight)
        ctx.geom, ctx.mode = geom, mode
        return out

    @staticmethod
    def backward(ctx, d_out):
        src, weight = ctx.saved_tensors
        geom, mode = ctx.geom, ctx.mode
        d_out = d_out.contiguous().float()
        co, ci = weight.shape[0], weight.shape[-1]
        d_src = d_w = None
        if ctx.needs_input_grad[0]:
            wp, m = pack_param(weight, True, geom.mirror, mode)
            d_src = sparse.conv([Seg(d_out, wp, geom.index_t, geom.mask_t)], geom.n_in, ci, m, raw=True)
        if ctx.needs_input_grad[1]:
            n_off = w_noff(weight)
            dw = torch.empty((n_off, ci, co), dtype=torch.float32, device=src.device)
            lib = _lib.load()
            idx = geom.index
            istride = 0 if idx is None else idx.stride(0)
            if mode != _lib.MODE_FP32 and USE_WGRAD_TC and lib.tl_conv_wgrad_tc_eligible(ci, co):
                # tcgen05 weight gradient (csrc/tl_wgrad_tc.cu): TF32 operands straight from the fp32 tensors
                wsb = lib.tl_conv_wgrad_tc_workspace_bytes(geom.n_out, ci, n_off, co)
                ws = torch.empty(max(int(wsb), 256), dtype=torch.uint8, device=src.device)
                check(lib.tl_conv_wgrad_tc(ptr(src), src.stride(0), ci, n_off, ptr(idx), istride, ptr(geom.mask), ptr(d_out),
                                           geom.n_out, co, ptr(dw), ptr(ws), wsb, stream_ptr()))
            else:
                check(lib.tl_conv_wgrad(ptr(src), src.stride(0), ci, n_off, ptr(idx), istride, ptr(geom.mask), ptr(d_out),
                                        geom.n_out, co, ptr(dw), int(mode != _lib.MODE_FP32), stream_ptr()))
            d_w = dw.permute(2, 0, 1).reshape(weight.shape).to(weight.dtype)
        return d_src, d_w, None, None


def w_noff(weight):
    return weight.shape[1] * weight.shape[2] * weight.shape[3]


def sparse_conv(src, weight, geom, mode):
    """src [n_in, C_in] -> [n_out, C_out]; weight in spconv's KRSC layout [C_out, k, k, k, C_in]."""
    if geom.n_out == 0 or geom.n_in == 0:
        return src.new_zeros((geom.n_out, weight.shape[0])) + 0 * weight.sum()
    return _SparseConv.apply(src, weight, geom, mode)


class _BNReLU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, eps, momentum, batch_stats):
        lib = _lib.load()
        x = x.contiguous().float()
        n, c = x.shape
        dev = x.device
        stats = torch.empty((4, c), dtype=torch.float32, device=dev)   # mean, invstd, scale, shift
        acc = torch.empty(2 * c, dtype=torch.float64, device=dev)
        g, b = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        if batch_stats:
            check(lib.tl_bn_stats(ptr(x), n, c, ptr(acc), stream_ptr()))
            check(lib.tl_bn_finalize(ptr(acc), n, c, ptr(g), ptr(b), float(eps), float(momentum), ptr(running_mean),
                                     ptr(running_var), ptr(stats[0]), ptr(stats[1]), ptr(stats[2]), ptr(stats[3]),
                                     stream_ptr()))
        else:
            stats[0] = running_mean
            stats[1] = torch.rsqrt(running_var.float() + eps)
            stats[2] = g * stats[1]
            stats[3] = b - stats[0] * stats[2]
        out = torch.empty_like(x)
        check(lib.tl_bn_relu_apply(ptr(x), n, c, ptr(stats[2]), ptr(stats[3]), ptr(out), stream_ptr()))
        ctx.save_for_backward(x, stats)
        ctx.batch_stats = bool(batch_stats)
        ctx.acc = acc
        return out

    @staticmethod
    def backward(ctx, d_act):
        lib = _lib.load()
        x, stats = ctx.saved_tensors
        n, c = x.shape
        d_act = d_act.contiguous().float()
        dx = torch.empty_like(x)
        dgb = torch.empty((2, c), dtype=torch.float32, device=x.device)
        check(lib.tl_bn_relu_bwd(ptr(x), ptr(d_act), n, c, ptr(stats[2]), ptr(stats[3]), ptr(stats[0]), ptr(stats[1]),
                                 int(ctx.batch_stats), ptr(ctx.acc), ptr(dx), ptr(dgb[0]), ptr(dgb[1]), stream_ptr()))
        return (dx if ctx.needs_input_grad[0] else None, dgb[0] if ctx.needs_input_grad[1] else None,
                dgb[1] if ctx.needs_input_grad[2] else None, None, None, None, None, None)


def bn_relu(x, bn):
    """relu(BatchNorm1d(x)) on feature rows; batch statistics iff `bn.training` (fixed_modules keep BN in eval,
    reference tree_learn.py:66-72).  Skipped for 0 rows like spconv's SparseSequential."""
    if x.shape[0] == 0:
        return x
    batch_stats = bn.training
    if batch_stats:
        if x.shape[0] == 1:
            raise ValueError(f'Expected more than 1 value per channel when training, got input size {list(x.shape)}')
        bn.num_batches_tracked += 1
    return _BNReLU.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps, bn.momentum, batch_stats)
