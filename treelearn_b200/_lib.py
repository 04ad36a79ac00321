"""ctypes binding of the C-ABI library (include/treelearn_b200.h).

The product path has NO CPU / eager fallback: if `libtreelearn_b200.so` is missing or does not
export a declared symbol this module raises at import of the first op (loudly), as the contract asks.
Build it with `python __graft_entry__.py` (or `make -C treelearn_b200/csrc`).
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('TL_LIB') or os.path.join(_HERE, 'libtreelearn_b200.so')   # TL_LIB: e.g. the TRACE=1 build

TL_MAX_SEG = 3
TILE_ROWS = 128
MODE_FP32, MODE_TF32, MODE_F16, MODE_F16X2 = 0, 1, 2, 3
ERR_REACH_ZERO = -3


class ConvSeg(C.Structure):
    _fields_ = [('src', C.c_void_p), ('src_stride', C.c_int64), ('c_in', C.c_int32), ('n_off', C.c_int32),
                ('index', C.c_void_p), ('index_stride', C.c_int64), ('tile_mask', C.c_void_p),
                ('weight', C.c_void_p)]


class ConvDesc(C.Structure):
    _fields_ = [('n_out', C.c_int32), ('c_out', C.c_int32), ('n_seg', C.c_int32), ('src_fp32_mask', C.c_int32),
                ('seg', ConvSeg * TL_MAX_SEG), ('residual', C.c_void_p), ('out_raw', C.c_void_p),
                ('out_act1', C.c_void_p), ('scale1', C.c_void_p), ('shift1', C.c_void_p),
                ('out_act2', C.c_void_p), ('scale2', C.c_void_p), ('shift2', C.c_void_p), ('splitk_ws', C.c_void_p),
                ('halo_rows', C.c_void_p), ('halo_cnt', C.c_void_p), ('halo_lidx', C.c_void_p), ('halo_cap', C.c_int32),
                ('halo_umax', C.c_int32)]


_P, _I32, _I64, _F, _D, _SZ = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double, C.c_size_t
_I64P = C.POINTER(C.c_int64)
_I32P = C.POINTER(C.c_int32)

# name -> (restype, argtypes); must list every symbol include/treelearn_b200.h declares
SIGNATURES = {
    'tl_last_error': (C.c_char_p, []),
    'tl_version': (C.c_int, []),
    'tl_launch_count': (C.c_longlong, []),
    'tl_reset_launch_count': (None, []),
    'tl_voxelize_workspace_bytes': (_SZ, [_I64]),
    'tl_voxelize': (C.c_int, [_P, _P, _I32, _P, _I64, _I32, _F, _I32, _I32, _I32, _P, _P, _P, _P, _I64P, _P, _SZ, _P]),
    'tl_level_workspace_bytes': (_SZ, [_I64]),
    'tl_build_level': (C.c_int, [_P, _I64, _I32P, _P, _P, _P, _P, _P, _P, _I32P, _I64P, _P, _SZ, _P]),
    'tl_rulebook_workspace_bytes': (_SZ, [_I64]),
    'tl_subm_rulebook': (C.c_int, [_P, _I64, _I32P, _P, _P, _P, _SZ, _P]),
    'tl_halo_build': (C.c_int, [_P, _P, _I64, _I64, _I32, _P, _P, _P, _P, _P]),
    'tl_conv_fwd': (C.c_int, [C.POINTER(ConvDesc), _I32, _P]),
    'tl_heads_fwd': (C.c_int, [_P, _I32, _P, _I64, _I32] + [_P] * 8 + [_P, _P, _P, _P]),
    'tl_merge_workspace_bytes': (_SZ, [_I64]),
    'tl_merge_groupby_mean': (C.c_int, [_P, _P, _I64, _I32, _P, _P, _P, _I64P, _P, _SZ, _P]),
    'tl_cluster_workspace_bytes': (_SZ, [_I64]),
    'tl_cluster_radius_cc': (C.c_int, [_P, _I64, _D, _I64, _I64, _I64, _P, _I64P, _P, _SZ, _P]),
    'tl_knn_workspace_bytes': (_SZ, [_I64, _I64]),
    'tl_knn_vote': (C.c_int, [_P, _P, _I64, _P, _I64, _I32, _P, _P, _SZ, _P]),
    'tl_hdbscan_workspace_bytes': (_SZ, [_I64]),
    'tl_core_distance': (C.c_int, [_P, _I64, _I32, _P, _P, _SZ, _P]),
    'tl_mst_prim': (C.c_int, [_P, _P, _I64, _P, _P, _P, _P, _SZ, _P]),
    'tl_hdbscan_tree_labels': (C.c_int, [_P, _P, _P, _I64, _I64, _P]),
    'tl_bn_stats': (C.c_int, [_P, _I64, _I32, _P, _P]),
    'tl_bn_finalize': (C.c_int, [_P, _I64, _I32, _P, _P, _F, _F, _P, _P, _P, _P, _P, _P, _P]),
    'tl_bn_relu_apply': (C.c_int, [_P, _I64, _I32, _P, _P, _P, _P]),
    'tl_bn_relu_bwd': (C.c_int, [_P, _P, _I64, _I32, _P, _P, _P, _P, _I32, _P, _P, _P, _P, _P]),
    'tl_pack_weight_tc': (C.c_int, [_P, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _P, _P]),
    'tl_conv_wgrad': (C.c_int, [_P, _I64, _I32, _I32, _P, _I64, _P, _P, _I64, _I32, _P, _I32, _P]),
    'tl_conv_wgrad_tc_eligible': (C.c_int, [_I32, _I32]),
    'tl_conv_wgrad_tc_workspace_bytes': (_SZ, [_I64, _I32, _I32, _I32]),
    'tl_conv_wgrad_tc': (C.c_int, [_P, _I64, _I32, _I32, _P, _I64, _P, _P, _I64, _I32, _P, _P, _SZ, _P]),
    'tl_hash_join_workspace_bytes': (_SZ, [_I64]),
    'tl_hash_join_last': (C.c_int, [_P, _I32, _I32, _P, _I64, _P, _I32, _I32, _I64, _I64, _P, _P, _SZ, _P]),
    'tl_cooccurrence_counts': (C.c_int, [_P, _P, _I64, _I64, _I64, _P, _P]),
    'tl_downsample_workspace_bytes': (_SZ, [_I64]),
    'tl_voxel_downsample_trace': (C.c_int, [_P, _I64, _I32, _D, _D, _P, _P, _P, _P, _I64P, _P, _SZ, _P]),
    'tl_verticality_workspace_bytes': (_SZ, [_I64]),
    'tl_verticality': (C.c_int, [_P, _I64, _D, _P, _P, _SZ, _P]),
}

_lib = None


def conv_source_hash():
    """sha1 over the sparse-conv kernel sources: profiles/*_conv_traffic.json records it so that bench.py can tell whether
    a committed ncu traffic capture still describes the kernels it is timing."""
    import hashlib
    h = hashlib.sha1()
    for f in ('tl_conv_halo.cu', 'tl_conv_grp.cu', 'tl_conv_ts.cu', 'tl_conv_tc.cu', 'tl_conv_simt.cu', 'tl_tc_ptx.cuh'):
        with open(os.path.join(_HERE, 'csrc', f), 'rb') as fh:
            h.update(fh.read())
    return h.hexdigest()


def load():
    """Load the shared library (once) and bind every declared symbol.  Raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f'{LIB_PATH} is missing: the CUDA extension was not built '
                           f'(run `python __graft_entry__.py` / `make -C treelearn_b200/csrc`). '
                           f'treelearn_b200 has no CPU fallback.')
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


class TreeLearnCudaError(RuntimeError):
    pass


def check(rc):
    if rc == 0:
        return
    msg = load().tl_last_error().decode()
    if rc == ERR_REACH_ZERO:
        raise ValueError(msg)   # message contains "reach zero!!!" (tree_learn/util/pipeline.py:91-97)
    raise TreeLearnCudaError(f'treelearn_b200 error {rc}: {msg}')


def ptr(t):
    return None if t is None else t.data_ptr()


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise TreeLearnCudaError('treelearn_b200 ops need CUDA tensors (no CPU path exists)')
