"""ORACLE (test infrastructure, never shipped, never measured as the product).

CPU restatement of the subset of the third-party `spconv` 2.x operator library that
TreeLearn calls.  `spconv` (setup/requirements.txt:1 `spconv-cu118`, version unpinned,
2.x line) is NOT vendored under /root/reference and is not installable here, so this
file restates its published semantics; the reference holds no test or golden vector at
this boundary => **parity unpinned** at the spconv boundary.

Call sites restated (reference file:line):
  * SparseConvTensor            tree_learn/model/tree_learn.py:88, blocks.py:35-38,73-77,140-147
  * SparseSequential            tree_learn/model/tree_learn.py:37-42, blocks.py:49-70,97-135
  * SubMConv3d                  tree_learn/model/tree_learn.py:37-39, blocks.py:57-70
  * SparseConv3d (k2,s2 / k1)   blocks.py:29-39,104-110
  * SparseInverseConv3d         blocks.py:118-123
  * PointToVoxel                tree_learn/model/tree_learn.py:136-143

Two independent implementations of each conv exist here so that the oracle checks itself:
  * rulebook form  (gather - GEMM - scatter with explicit neighbour tables; any size)
  * dense form     (`dense_*` helpers: F.conv3d / F.conv_transpose3d masked to the active set;
                    exact by construction, only for small grids)
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

_COORD_BITS = 16  # oracle keys never alias (spconv's own linearised keys can)


# --------------------------------------------------------------------------------------
# integer helpers (numpy, int64) -- the bit-exact part of the oracle
# --------------------------------------------------------------------------------------
def pack_keys(indices):
    """(b,x,y,z) int rows -> one int64 key per row (no aliasing for coords < 2**16)."""
    idx = np.asarray(indices).astype(np.int64)
    return (((idx[:, 0] << _COORD_BITS | idx[:, 1]) << _COORD_BITS | idx[:, 2]) << _COORD_BITS) | idx[:, 3]


def subm_neighbour_table(indices, spatial_shape, kernel_size=3):
    """nbr[k, j] = row of the active voxel at coords(j) + (kappa - pad) in the same batch, or -1.

    k enumerates kappa in C order over (k0,k1,k2) which follow the index columns (x,y,z)
    [spconv-knowledge, SURVEY App. A.3].  Out-of-`spatial_shape` neighbours do not exist.
    """
    idx = np.asarray(indices).astype(np.int64)
    n = idx.shape[0]
    ks = int(kernel_size)
    pad = (ks - 1) // 2
    nbr = -np.ones((ks ** 3, n), dtype=np.int64)
    if n == 0:
        return nbr
    keys = pack_keys(idx)
    order = np.argsort(keys, kind='stable')
    skeys = keys[order]
    shape = np.asarray(spatial_shape, dtype=np.int64).reshape(3)
    k = 0
    for a in range(ks):
        for b in range(ks):
            for c in range(ks):
                q = idx.copy()
                q[:, 1] += a - pad
                q[:, 2] += b - pad
                q[:, 3] += c - pad
                ok = np.all((q[:, 1:] >= 0) & (q[:, 1:] < shape[None, :]), axis=1)
                q[~ok, 1:] = 0
                qk = pack_keys(q)
                pos = np.searchsorted(skeys, qk)
                pos[pos >= n] = n - 1
                hit = ok & (skeys[pos] == qk)
                nbr[k, hit] = order[pos[hit]]
                k += 1
    return nbr


def strided_pairs(indices, spatial_shape, kernel_size=2, stride=2):
    """SparseConv3d(k=2,s=2,pad=0) rulebook.

    Returns (out_indices [m,4] int32, out_shape[3], in_row [P], kappa [P], out_row [P]).
    out spatial = floor((S-k)/s)+1; a pair (p,kappa) exists iff (p-kappa) % s == 0 and
    q=(p-kappa)/s lies in [0,out_shape); output rows = sorted unique q (the GPU order of
    spconv; order is implementation defined and cancels out) [spconv-knowledge App. A.3].
    Raises ValueError containing 'reach zero!!!' when an output axis collapses
    (caught by tree_learn/util/pipeline.py:91-97).
    """
    assert kernel_size == 2 and stride == 2, 'TreeLearn only uses k=2,s=2 strided convs'
    idx = np.asarray(indices).astype(np.int64)
    shape = [int(s) for s in np.asarray(spatial_shape).reshape(3)]
    out_shape = [(s - kernel_size) // stride + 1 for s in shape]
    if min(out_shape) <= 0:
        raise ValueError(f'your out spatial shape {out_shape} reach zero!!! input shape: {shape}')
    q = idx.copy()
    q[:, 1:] = idx[:, 1:] // 2
    kap = idx[:, 1:] - 2 * q[:, 1:]
    ok = np.all(q[:, 1:] < np.asarray(out_shape)[None, :], axis=1)
    in_row = np.nonzero(ok)[0]
    qk = pack_keys(q[ok])
    uq, inv = np.unique(qk, return_inverse=True)
    out_idx = np.stack([uq >> (3 * _COORD_BITS), (uq >> (2 * _COORD_BITS)) & 0xFFFF,
                        (uq >> _COORD_BITS) & 0xFFFF, uq & 0xFFFF], axis=1).astype(np.int32)
    kappa = (kap[ok, 0] * 4 + kap[ok, 1] * 2 + kap[ok, 2]).astype(np.int64)
    return out_idx, out_shape, in_row, kappa, inv.astype(np.int64)


# --------------------------------------------------------------------------------------
# container + module plumbing
# --------------------------------------------------------------------------------------
class SparseConvTensor:
    def __init__(self, features, indices, spatial_shape, batch_size, grid=None):
        self._features = features
        self.indices = indices
        if torch.is_tensor(spatial_shape):  # the reference passes a tensor (tree_learn.py:86-88)
            spatial_shape = [int(v) for v in spatial_shape.tolist()]
        self.spatial_shape = [int(v) for v in spatial_shape]
        self.batch_size = batch_size
        self.indice_dict = {}
        self.grid = grid

    @property
    def features(self):
        return self._features

    @features.setter
    def features(self, _):
        raise ValueError('features is read-only; use replace_feature')

    def replace_feature(self, feature):
        out = SparseConvTensor(feature, self.indices, self.spatial_shape, self.batch_size, self.grid)
        out.indice_dict = self.indice_dict
        return out


class SparseModule(nn.Module):
    pass


class SparseSequential(SparseModule):
    def __init__(self, *args, **kwargs):
        super().__init__()
        if len(args) == 1 and isinstance(args[0], OrderedDict):
            for key, module in args[0].items():
                self.add_module(key, module)
        else:
            for i, module in enumerate(args):
                self.add_module(str(i), module)
        for name, module in kwargs.items():
            self.add_module(name, module)

    def forward(self, x):
        for module in self._modules.values():
            if isinstance(module, SparseModule):
                x = module(x)
            elif isinstance(x, SparseConvTensor):
                if x.indices.shape[0] != 0:
                    x = x.replace_feature(module(x.features))
            else:
                x = module(x)
        return x


class _ConvBase(SparseModule):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, bias=True, indice_key=None, algo=None, **kwargs):
        super().__init__()
        assert dilation == 1 and groups == 1
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.padding = int(kernel_size), int(stride), int(padding)
        self.indice_key = indice_key
        k = self.kernel_size
        self.weight = nn.Parameter(torch.empty(out_channels, k, k, k, in_channels))  # KRSC
        nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None

    def _finish(self, feats):
        return feats if self.bias is None else feats + self.bias


def _gather_gemm_scatter(feats, weight, in_rows, kappas, out_rows, n_out):
    """out[o] += W[:,kappa,:] @ in[i] for every pair; fp32 GEMMs, index_add in pair order."""
    co = weight.shape[0]
    w = weight.reshape(co, -1, weight.shape[-1])
    out = feats.new_zeros((n_out, co))
    in_rows = torch.as_tensor(in_rows)
    kappas = torch.as_tensor(kappas)
    out_rows = torch.as_tensor(out_rows)
    for k in range(w.shape[1]):
        sel = kappas == k
        if bool(sel.any()):
            out.index_add_(0, out_rows[sel], feats[in_rows[sel]] @ w[:, k, :].T)
    return out


class SubMConv3d(_ConvBase):
    def forward(self, x):
        key = self.indice_key
        nbr = x.indice_dict.get(key) if key is not None else None
        if nbr is None:
            nbr = subm_neighbour_table(x.indices.cpu().numpy(), x.spatial_shape, self.kernel_size)
            if key is not None:
                x.indice_dict[key] = nbr
        n = x.features.shape[0]
        co = self.out_channels
        w = self.weight.reshape(co, -1, self.in_channels)
        out = x.features.new_zeros((n, co))
        for k in range(nbr.shape[0]):
            dst = np.nonzero(nbr[k] >= 0)[0]
            if dst.size:
                src = torch.from_numpy(nbr[k][dst])
                out.index_add_(0, torch.from_numpy(dst), x.features[src] @ w[:, k, :].T)
        return x.replace_feature(self._finish(out))


class SparseConv3d(_ConvBase):
    def forward(self, x):
        out_idx, out_shape, in_row, kappa, out_row = strided_pairs(
            x.indices.cpu().numpy(), x.spatial_shape, self.kernel_size, self.stride)
        if self.indice_key is not None:
            x.indice_dict[self.indice_key] = dict(in_indices=x.indices, in_shape=x.spatial_shape,
                                                  in_row=in_row, kappa=kappa, out_row=out_row)
        feats = _gather_gemm_scatter(x.features, self.weight, in_row, kappa, out_row, out_idx.shape[0])
        out = SparseConvTensor(self._finish(feats), torch.from_numpy(out_idx), out_shape, x.batch_size, x.grid)
        out.indice_dict = x.indice_dict
        return out


class SparseInverseConv3d(_ConvBase):
    def __init__(self, in_channels, out_channels, kernel_size, indice_key=None, bias=True, **kwargs):
        super().__init__(in_channels, out_channels, kernel_size, bias=bias, indice_key=indice_key)

    def forward(self, x):
        saved = x.indice_dict[self.indice_key]
        n_fine = saved['in_indices'].shape[0]
        # pairs swapped: fine row <- coarse row, same kappa, no kernel flip (App. A.3)
        feats = _gather_gemm_scatter(x.features, self.weight, saved['out_row'], saved['kappa'],
                                     saved['in_row'], n_fine)
        out = SparseConvTensor(self._finish(feats), saved['in_indices'], saved['in_shape'], x.batch_size, x.grid)
        out.indice_dict = x.indice_dict
        return out


# --------------------------------------------------------------------------------------
# point -> voxel (CPU semantics of spconv's Point2Voxel: first-seen ids, first <=P points)
# --------------------------------------------------------------------------------------
def voxel_index_fp32(points_xyz, range_min, vsize):
    """c = floor((p - min) / vsize) evaluated in float32 like spconv's kernel (App. A.4)."""
    p = np.asarray(points_xyz, dtype=np.float32)
    mn = np.asarray(range_min, dtype=np.float32)
    vs = np.asarray(vsize, dtype=np.float32)
    return np.floor((p - mn[None, :]) / vs[None, :]).astype(np.int64)


class PointToVoxel:
    def __init__(self, vsize_xyz, coors_range_xyz, num_point_features, max_num_voxels,
                 max_num_points_per_voxel, device=None):
        self.vsize = [float(v) for v in vsize_xyz]
        self.range = [float(v) for v in coors_range_xyz]
        self.num_point_features = num_point_features
        self.max_num_voxels = max_num_voxels
        self.max_pts = max_num_points_per_voxel
        self.grid_size = [int(round((self.range[3 + i] - self.range[i]) / self.vsize[i])) for i in range(3)]

    def generate_voxel_with_id(self, pc):
        pts = pc.detach().cpu().numpy().astype(np.float32)
        n, f = pts.shape
        c = voxel_index_fp32(pts[:, :3], self.range[:3], self.vsize)
        inside = np.all((c >= 0) & (c < np.asarray(self.grid_size)[None, :]), axis=1)
        lin = (c[:, 0] * self.grid_size[1] + c[:, 1]) * self.grid_size[2] + c[:, 2]
        lin[~inside] = -1
        uq, first, inv = np.unique(lin[inside], return_index=True, return_inverse=True)
        rank = np.empty(uq.shape[0], dtype=np.int64)
        rank[np.argsort(first, kind='stable')] = np.arange(uq.shape[0])  # first-seen order
        vid_inside = rank[inv]
        m = min(uq.shape[0], self.max_num_voxels)
        pc_voxel_id = -np.ones(n, dtype=np.int64)
        pc_voxel_id[inside] = np.where(vid_inside < m, vid_inside, -1)
        voxels = np.zeros((m, self.max_pts, f), dtype=np.float32)
        num = np.zeros(m, dtype=np.int32)
        coords = np.zeros((m, 3), dtype=np.int32)
        for j in np.nonzero(pc_voxel_id >= 0)[0]:
            v = pc_voxel_id[j]
            if num[v] == 0:
                coords[v] = c[j, ::-1]  # zyx
            if num[v] < self.max_pts:
                voxels[v, num[v]] = pts[j]
                num[v] += 1
        dev = pc.device
        return (torch.from_numpy(voxels).to(dev), torch.from_numpy(coords).to(dev),
                torch.from_numpy(num).to(dev), torch.from_numpy(pc_voxel_id).to(dev))


# --------------------------------------------------------------------------------------
# dense cross-check forms (exact; small grids only)
# --------------------------------------------------------------------------------------
def _densify(feats, indices, spatial_shape, batch_size):
    c = feats.shape[1]
    dense = feats.new_zeros((batch_size, c, *spatial_shape))
    idx = indices.long()
    dense[idx[:, 0], :, idx[:, 1], idx[:, 2], idx[:, 3]] = feats
    return dense


def _sample(dense, indices):
    idx = indices.long()
    return dense[idx[:, 0], :, idx[:, 1], idx[:, 2], idx[:, 3]]


def dense_subm(feats, indices, spatial_shape, batch_size, weight):
    d = _densify(feats, indices, spatial_shape, batch_size)
    pad = (weight.shape[1] - 1) // 2
    return _sample(F.conv3d(d, weight.permute(0, 4, 1, 2, 3), padding=pad), indices)


def dense_strided(feats, indices, spatial_shape, batch_size, weight, out_indices):
    d = _densify(feats, indices, spatial_shape, batch_size)
    return _sample(F.conv3d(d, weight.permute(0, 4, 1, 2, 3), stride=2), out_indices)


def dense_inverse(coarse_feats, coarse_indices, coarse_shape, batch_size, weight, fine_indices, fine_shape):
    d = _densify(coarse_feats, coarse_indices, coarse_shape, batch_size)
    up = F.conv_transpose3d(d, weight.permute(4, 0, 1, 2, 3), stride=2)
    full = up.new_zeros((batch_size, up.shape[1], *fine_shape))
    sx, sy, sz = (min(a, b) for a, b in zip(up.shape[2:], fine_shape))
    full[:, :, :sx, :sy, :sz] = up[:, :, :sx, :sy, :sz]
    return _sample(full, fine_indices)


def install_as_spconv():
    """Expose this module as `spconv.pytorch[.utils|.modules]` so the reference's own
    tree_learn/model/{tree_learn,blocks}.py can be imported verbatim (fixture generation only)."""
    import sys
    import types
    this = sys.modules[__name__]
    root = types.ModuleType('spconv')
    pt = types.ModuleType('spconv.pytorch')
    for name in ('SparseConvTensor', 'SparseSequential', 'SubMConv3d', 'SparseConv3d',
                 'SparseInverseConv3d', 'SparseModule'):
        setattr(pt, name, getattr(this, name))
    utils = types.ModuleType('spconv.pytorch.utils')
    utils.PointToVoxel = PointToVoxel
    mods = types.ModuleType('spconv.pytorch.modules')
    mods.SparseModule = SparseModule
    pt.utils, pt.modules, root.pytorch = utils, mods, pt
    sys.modules.update({'spconv': root, 'spconv.pytorch': pt, 'spconv.pytorch.utils': utils,
                        'spconv.pytorch.modules': mods})
