"""The step before the per-tile path, same names / arguments as the reference, numpy in and numpy
out, computed on the GPU and kept in memory (no .npz round trip between the stages):

  voxelize           <- tree_learn/util/data_preparation.py:60-79     (open3d voxel_down_sample_and_trace -> tl_voxel_downsample_trace)
  compute_features   <- tree_learn/util/data_preparation.py:83-100    (jakteristics verticality -> tl_verticality, NaN -> column mean)
  tile_grid / cut_tiles <- data_preparation.py:333-494                (`SampleGenerator.tile_generate_and_save`, default settings:
                                                                        plot_corners=None, no denoising) -> list of tile dicts
  prepare_tiles      <- tree_learn/util/pipeline.py:24-75             (`generate_tiles` without the files)

open3d / jakteristics are not in this image: the first two follow the libraries' published algorithms (parity unpinned,
see oracle/prepare_ref.py); tile cutting is pinned to the reference's own class by tests/golden/tiles_small.npz.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr
from .pipeline import _dev, _ws


class Trace:
    """Input indices per voxel (CSR), iterable like the list of index vectors open3d returns: `trace[v]` is an int64
    array in input order, `trace[v][0]` the first point that fell into voxel v."""

    def __init__(self, offsets, indices):
        self.offsets, self.indices = offsets, indices

    def __len__(self):
        return len(self.offsets) - 1

    def __getitem__(self, v):
        if not -len(self) <= v < len(self):
            raise IndexError(v)
        v %= len(self)
        return self.indices[self.offsets[v]:self.offsets[v + 1]]

    def __iter__(self):
        return (self[v] for v in range(len(self)))


def voxel_downsample_trace_cuda(points, voxel_size, voxel_min_bound, round2_first=False):
    """points [n,3] f64 CUDA -> (voxel points [m,3] f64, first_index [m] i64, offsets [m+1] i64, trace [n] i64)."""
    lib = _lib.load()
    assert points.dtype == torch.float64 and points.is_contiguous() and points.shape[1] == 3
    n, dev = int(points.shape[0]), points.device
    out = torch.empty((n, 3), dtype=torch.float64, device=dev)
    first = torch.empty(n, dtype=torch.int64, device=dev)
    offsets = torch.empty(n + 1, dtype=torch.int64, device=dev)
    trace = torch.empty(n, dtype=torch.int64, device=dev)
    wsb = lib.tl_downsample_workspace_bytes(n)
    ws = _ws(wsb, dev)
    m = C.c_int64(0)
    check(lib.tl_voxel_downsample_trace(ptr(points), n, int(round2_first), float(voxel_size), float(voxel_min_bound),
                                        ptr(out), ptr(first), ptr(offsets), ptr(trace), C.byref(m), ptr(ws), wsb,
                                        stream_ptr()))
    m = int(m.value)
    return out[:m], first[:m], offsets[:m + 1], trace


def voxelize(data, voxel_size):
    """Down-sample to one point per voxel (the mean of its points rounded to 2 decimals); the extra columns of a voxel
    come from its first point.  Returns (data [m, C] float64, Trace).  Voxels come in ascending grid order (open3d's
    order is that of its hash map, i.e. unspecified)."""
    dev = _dev()
    data = np.asarray(data)
    pts = torch.from_numpy(np.ascontiguousarray(data[:, :3], dtype=np.float64)).to(dev)
    if pts.shape[0] == 0:
        return np.zeros((0, data.shape[1])), Trace(np.zeros(1, np.int64), np.zeros(0, np.int64))
    rounded_abs_max = float((torch.round(pts * 100.0) / 100.0).abs().max().item())      # np.max(np.abs(np.round(points, 2)))
    bound = rounded_abs_max + 100
    down, first, offsets, trace = voxel_downsample_trace_cuda(pts, voxel_size, -bound - voxel_size * 0.5, round2_first=True)
    down, first = down.cpu().numpy(), first.cpu().numpy()
    if data.shape[1] >= 4:
        down = np.hstack((down, data[:, 3:][first]))
    return down, Trace(offsets.cpu().numpy(), trace.cpu().numpy())


def verticality_cuda(points, search_radius):
    """points [n,3] f64 CUDA -> [n] f64 (NaN where fewer than 3 points lie within the radius)."""
    lib = _lib.load()
    assert points.dtype == torch.float64 and points.is_contiguous() and points.shape[1] == 3
    n = int(points.shape[0])
    out = torch.empty(n, dtype=torch.float64, device=points.device)
    wsb = lib.tl_verticality_workspace_bytes(n)
    ws = _ws(wsb, points.device)
    check(lib.tl_verticality(ptr(points), n, float(search_radius), ptr(out), ptr(ws), wsb, stream_ptr()))
    return out


def compute_features(points, search_radius=0.6, feature_names=['verticality'], num_threads=4):
    """[n,1] float32 verticality, NaNs replaced by the mean of the valid values (`replace_nanfeatures`)."""
    if list(feature_names) != ['verticality']:
        raise NotImplementedError('only the verticality feature (the one the reference feeds to the network) is built')
    points = np.asarray(points)
    assert points.shape[1] == 3
    dev = _dev()
    v = verticality_cuda(torch.from_numpy(np.ascontiguousarray(points, dtype=np.float64)).to(dev), search_radius)
    nan = torch.isnan(v)
    print(f'There are {int(nan.sum())} nan features in the whole forest. Replacing them with mean feature values')
    if nan.any():
        v[nan] = v[~nan].mean() if (~nan).any() else float('nan')
    return v.cpu().numpy().astype(np.float32).reshape(-1, 1)


def tile_grid(x_range, y_range, inner_edge, outer_edge, stride):
    """Inner squares [T,4] = (xmin, xmax, ymin, ymax) of the tile grid, row by row from the top; every scalar operation
    is done in the type the reference does it in (x_range / y_range are float32 when the plot was stored as float32)."""
    lo_x, hi_x = np.round(x_range[0] - 1.5 * outer_edge, 2), np.round(x_range[1] + 1.5 * outer_edge, 2)
    lo_y, hi_y = np.round(y_range[0] - 1.5 * outer_edge, 2), np.round(y_range[1] + 1.5 * outer_edge, 2)

    def axis(lo, hi):
        count = int(np.round((hi - lo - 2 * outer_edge) / inner_edge))
        edge = np.round((hi - lo - 2 * outer_edge) / count, 5)
        return int((count - 1) / stride + 1), edge

    ncols, edge_x = axis(lo_x, hi_x)
    nrows, edge_y = axis(lo_y, hi_y)
    cols = [(lo_x + outer_edge + stride * j * edge_x, lo_x + outer_edge + (stride * j + 1) * edge_x) for j in range(ncols)]
    rows = [(hi_y - outer_edge - (stride * i + 1) * edge_y, hi_y - outer_edge - stride * i * edge_y) for i in range(nrows)]
    squares = np.empty((nrows * ncols, 4))
    for i, (y0, y1) in enumerate(rows):
        for j, (x0, x1) in enumerate(cols):
            squares[i * ncols + j] = (x0, x1, y0, y1)
    return np.round(squares, 5)


def cut_tiles(points, labels, feats, inner_edge, outer_edge, stride, device=None):
    """Overlapping tiles of a voxelised plot: points [n,3] f32, labels [n] f32, feats [n,F] f32 -> list of dicts with
    the arrays the reference saves per tile ('points' f32 centred, 'feat' f32, 'instance_label' i32, 'center' f64 [3]);
    tiles whose inner square holds no point are dropped.  `device` defaults to the current CUDA device."""
    dev = _dev() if device is None else torch.device(device)
    points, labels, feats = np.asarray(points), np.asarray(labels), np.asarray(feats)
    table = np.hstack([np.hstack((points, labels.reshape(-1, 1))), feats])
    x_range, y_range = (points[:, 0].min(), points[:, 0].max()), (points[:, 1].min(), points[:, 1].max())
    inner = tile_grid(x_range, y_range, inner_edge, outer_edge, stride)
    outer = inner + np.array([-outer_edge, outer_edge, -outer_edge, outer_edge]).reshape(1, 4)
    rows = torch.from_numpy(table).to(dev)
    x, y = rows[:, 0], rows[:, 1]
    x64, y64 = x.double(), y.double()
    outer_t, inner_t = torch.from_numpy(outer).to(dev).to(rows.dtype), torch.from_numpy(inner).to(dev)
    # the outer bounds are compared in the table's dtype (a 0-dim fp64 bound against an fp32 tensor is cast down by
    # torch), the inner bounds in fp64 (numpy compares an fp32 array with an fp64 scalar in fp64)
    tiles = []
    for t in range(len(inner)):
        o, q = outer_t[t], inner_t[t]
        sel = (x >= o[0]) & (x <= o[1]) & (y >= o[2]) & (y <= o[3])
        has_inner = sel & (x64 >= q[0]) & (x64 < q[1]) & (y64 > q[2]) & (y64 <= q[3])
        if not bool(has_inner.any()):
            continue
        sq32 = inner[t].astype(np.float32)
        center_x, center_y = np.round((sq32[0] + sq32[1]) / 2, 6), np.round((sq32[2] + sq32[3]) / 2, 6)
        shift = torch.zeros((1, rows.shape[1]), dtype=torch.float64, device=dev)
        shift[0, 0], shift[0, 1] = float(center_x), float(center_y)
        chunk = (rows[sel].double() - shift).float().cpu().numpy()
        tiles.append({'points': chunk[:, :3], 'feat': chunk[:, 4:], 'instance_label': chunk[:, 3].astype(np.int32),
                      'center': np.array([center_x, center_y, 0])})
    return tiles


def prepare_tiles(data, voxel_size=0.1, search_radius_features=0.6, inner_edge=8, outer_edge=13.5, stride=0.5):
    """`generate_tiles` in memory (tree_learn/util/pipeline.py:24-75): raw plot [N, >=4] (x, y, z, label) -> voxelised
    fp32 plot rounded to 2 decimals, its verticality, and the list of tiles.  Returns (plot [m,4] f32, features [m,1] f32,
    tiles, trace)."""
    down, trace = voxelize(data, voxel_size)
    plot = np.round(down.astype(np.float32), 2)
    feats = compute_features(plot[:, :3].astype(np.float64), search_radius_features)
    return plot, feats, cut_tiles(plot[:, :3], plot[:, 3], feats, inner_edge, outer_edge, stride), trace
