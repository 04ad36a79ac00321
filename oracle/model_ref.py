"""ORACLE (test infrastructure): functional CPU restatement of the TreeLearn network.

Walks a reference-layout `state_dict` with plain torch ops on top of the
restated rulebooks in `oracle/spconv_ref.py`.  Follows, by file:line of /root/reference:
  voxelisation       tree_learn/model/tree_learn.py:129-167   (`voxelize_ref`)
  backbone           tree_learn/model/tree_learn.py:83-94     (`backbone_ref`)
  residual block     tree_learn/model/blocks.py:42-79         (`_residual`)
  U-block recursion  tree_learn/model/blocks.py:81-149        (`_ublock`)
  heads              tree_learn/model/tree_learn.py:97-103, blocks.py:8-18 (`heads_ref`)
  loss               tree_learn/model/tree_learn.py:106-126, tree_learn/util/train.py:145-166 (`loss_ref`)
The structure is pinned against the reference's own model code (run on the restated spconv
namespace) by tests/golden/make_golden.py; the spconv boundary itself is **parity unpinned**.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/--impl reference may import it.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import spconv_ref as sp

BN_EPS = 1e-4        # tree_learn.py:34
BN_MOMENTUM = 0.1
LOSS_MULTIPLIER_SEMANTIC = 50  # tree_learn.py:9


def voxelize_ref(coords, input_feats, batch_ids, batch_size, voxel_size=0.1, use_coords=False,
                 use_feats=False, max_num_points_per_voxel=3, epsilon=1.0):
    """Point -> voxel mean pool.  Returns (voxel_feats [M,4] as [feat,x,y,z], voxel_indices [M,4]
    int32 (b,x,y,z), v2p [N] int64, spatial_shape[3]).  Voxel order: first seen, per batch."""
    feats = torch.hstack([coords, input_feats]).float()
    vf, vc, v2p = [], [], []
    total = 0
    for b in range(batch_size):
        sel = batch_ids == b
        pts = feats[sel].numpy()
        mn = pts[:, :3].min(axis=0)
        c = sp.voxel_index_fp32(pts[:, :3], mn, [voxel_size] * 3)
        lin = (c[:, 0] << 32) | (c[:, 1] << 16) | c[:, 2]
        uq, first, inv = np.unique(lin, return_index=True, return_inverse=True)
        rank = np.empty(len(uq), dtype=np.int64)
        rank[np.argsort(first, kind='stable')] = np.arange(len(uq))
        vid = rank[inv]
        m = len(uq)
        ssum = np.zeros((m, pts.shape[1]), dtype=np.float32)
        cnt = np.zeros(m, dtype=np.int64)
        kept = np.zeros(m, dtype=np.int64)
        vcoord = np.zeros((m, 3), dtype=np.int64)
        seen = np.zeros(m, dtype=bool)
        for j in range(len(pts)):
            v = vid[j]
            if not seen[v]:
                seen[v] = True
                vcoord[v] = c[j]
            if kept[v] < max_num_points_per_voxel:
                kept[v] += 1
                if not np.all(pts[j] == 0):      # all-zero rows are padding (tree_learn.py:149-150)
                    ssum[v] += pts[j]
                    cnt[v] += 1
        with np.errstate(invalid='ignore', divide='ignore'):
            mean = ssum / cnt[:, None].astype(np.float32)   # 0/0 -> NaN like nanmean of all-NaN
        if not use_coords:
            mean[:, :3] = 1.0
        if not use_feats:
            mean[:, 3:] = 1.0
        vf.append(np.concatenate([mean[:, 3:], mean[:, :3]], axis=1))
        vc.append(np.concatenate([np.full((m, 1), b, dtype=np.int64), vcoord], axis=1))
        v2p.append(vid + total)
        total += m
    vc = np.concatenate(vc)
    spatial_shape = (vc[:, 1:].max(axis=0) + 1).tolist()
    return (torch.from_numpy(np.concatenate(vf)).float(), torch.from_numpy(vc.astype(np.int32)),
            torch.from_numpy(np.concatenate(v2p)), spatial_shape)


class _Ctx:
    def __init__(self, sd, training, new_stats):
        self.sd, self.training, self.new_stats = sd, training, new_stats


def _bn_relu(ctx, prefix, x):
    sd = ctx.sd
    if x.shape[0] == 0:
        return x
    if ctx.training:
        rm, rv = sd[prefix + '.running_mean'].clone(), sd[prefix + '.running_var'].clone()
        y = F.batch_norm(x, rm, rv, sd[prefix + '.weight'], sd[prefix + '.bias'], True, BN_MOMENTUM, BN_EPS)
        ctx.new_stats[prefix + '.running_mean'], ctx.new_stats[prefix + '.running_var'] = rm, rv
    else:
        y = F.batch_norm(x, sd[prefix + '.running_mean'], sd[prefix + '.running_var'],
                         sd[prefix + '.weight'], sd[prefix + '.bias'], False, BN_MOMENTUM, BN_EPS)
    return F.relu(y)


def _subm(x, nbr, weight):
    co, ci = weight.shape[0], weight.shape[-1]
    w = weight.reshape(co, -1, ci)
    out = x.new_zeros((x.shape[0], co))
    for k in range(nbr.shape[0]):
        dst = np.nonzero(nbr[k] >= 0)[0]
        if dst.size:
            out = out.index_add(0, torch.from_numpy(dst), x[torch.from_numpy(nbr[k][dst])] @ w[:, k, :].T)
    return out


def _pairs_conv(x, weight, src, kappa, dst, n_out):
    co, ci = weight.shape[0], weight.shape[-1]
    w = weight.reshape(co, -1, ci)
    out = x.new_zeros((n_out, co))
    src, kappa, dst = (torch.as_tensor(a) for a in (src, kappa, dst))
    for k in range(w.shape[1]):
        sel = kappa == k
        if bool(sel.any()):
            out = out.index_add(0, dst[sel], x[src[sel]] @ w[:, k, :].T)
    return out


def _residual(ctx, p, x, nbr):
    sd = ctx.sd
    h = _subm(_bn_relu(ctx, p + '.conv_branch.0', x), nbr, sd[p + '.conv_branch.2.weight'])
    h = _subm(_bn_relu(ctx, p + '.conv_branch.3', h), nbr, sd[p + '.conv_branch.5.weight'])
    wi = sd.get(p + '.i_branch.0.weight')
    skip = x if wi is None else x @ wi.reshape(wi.shape[0], wi.shape[-1]).T
    return h + skip


def _ublock(ctx, p, x, indices, shape, levels_left):
    sd = ctx.sd
    nbr = sp.subm_neighbour_table(indices, shape, 3)
    for i in range(2):
        x = _residual(ctx, f'{p}.blocks.block{i}', x, nbr)
    if levels_left > 1:
        out_idx, out_shape, in_row, kappa, out_row = sp.strided_pairs(indices, shape)
        d = _pairs_conv(_bn_relu(ctx, p + '.conv.0', x), sd[p + '.conv.2.weight'], in_row, kappa, out_row, len(out_idx))
        d = _ublock(ctx, p + '.u', d, out_idx, out_shape, levels_left - 1)
        up = _pairs_conv(_bn_relu(ctx, p + '.deconv.0', d), sd[p + '.deconv.2.weight'], out_row, kappa, in_row, x.shape[0])
        x = torch.cat([x, up], dim=1)
        for i in range(2):
            x = _residual(ctx, f'{p}.blocks_tail.block{i}', x, nbr)
    return x


def backbone_ref(sd, voxel_feats, voxel_indices, spatial_shape, training=False, new_stats=None):
    """input_conv -> UBlock -> output_layer.  Returns [M,C] features (same row order as input)."""
    ctx = _Ctx(sd, training, {} if new_stats is None else new_stats)
    idx = voxel_indices.numpy() if torch.is_tensor(voxel_indices) else np.asarray(voxel_indices)
    num_levels = 1 + max([k.split('.').count('u') for k in sd if k.startswith('unet.')] + [0])
    nbr = sp.subm_neighbour_table(idx, spatial_shape, 3)
    x = _subm(voxel_feats, nbr, sd['input_conv.0.weight'])
    x = _ublock(ctx, 'unet', x, idx, spatial_shape, num_levels)
    return _bn_relu(ctx, 'output_layer.0', x)


def _mlp(ctx, p, x):
    sd = ctx.sd
    h = F.linear(x, sd[p + '.0.weight'], sd[p + '.0.bias'])
    h = _bn_relu(ctx, p + '.1', h)
    return F.linear(h, sd[p + '.3.weight'], sd[p + '.3.bias'])


def heads_ref(sd, voxel_out, v2p, training=False, new_stats=None):
    ctx = _Ctx(sd, training, {} if new_stats is None else new_stats)
    feats = voxel_out[v2p]
    return {'backbone_feats': feats,
            'semantic_prediction_logits': _mlp(ctx, 'semantic_linear', feats),
            'offset_predictions': _mlp(ctx, 'offset_linear', feats)}


def forward_ref(sd, batch, use_coords=False, use_feats=False, voxel_size=0.1, spatial_shape=None,
                max_num_points_per_voxel=3, training=False, new_stats=None):
    vf, vi, v2p, shape = voxelize_ref(batch['coords'], batch['input_feats'], batch['batch_ids'],
                                      batch['batch_size'], voxel_size, use_coords, use_feats,
                                      max_num_points_per_voxel)
    if spatial_shape is not None:
        shape = list(spatial_shape)
    out = backbone_ref(sd, vf, vi, shape, training, new_stats)
    return heads_ref(sd, out, v2p, training, new_stats)


def loss_ref(output, batch):
    """(loss, {'semantic_loss','offset_loss'}) -- tree_learn.py:106-126 + util/train.py:145-166."""
    logits = output['semantic_prediction_logits'].float()
    off = output['offset_predictions'].float()
    ms, mo = batch['masks_sem'], batch['masks_off']
    if int(ms.sum()) == 0:
        sem = 0 * logits.sum()
    else:
        sem = F.cross_entropy(logits[ms], batch['semantic_labels'][ms], reduction='sum') / int(ms.sum())
    if int(mo.sum()) == 0:
        offl = 0 * off.sum()
    else:
        offl = (off[mo] - batch['offset_labels'][mo]).pow(2).sum(1).sqrt().mean()
    d = {'semantic_loss': sem * LOSS_MULTIPLIER_SEMANTIC, 'offset_loss': offl}
    return d['semantic_loss'] + d['offset_loss'], d


# ---- deterministic weights with the reference's state_dict layout ----------
def make_state_dict(channels=32, num_blocks=7, dim_in=4, seed=0, randomize_bn=True):
    """Random weights in the reference layout; BN running stats randomised so BN is not a no-op
    (SURVEY §8d).  Conv weights ~ U(-b,b) with b = sqrt(1 / (27 * C_in)) (variance-preserving-ish)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(name, co, k, ci):
        b = (3.0 / (k ** 3 * ci)) ** 0.5
        sd[name] = (torch.rand((co, k, k, k, ci), generator=g) * 2 - 1) * b

    def bn(name, c):
        if randomize_bn:
            sd[name + '.weight'] = 0.5 + torch.rand(c, generator=g)
            sd[name + '.bias'] = 0.2 * torch.randn(c, generator=g)
            sd[name + '.running_mean'] = 0.1 * torch.randn(c, generator=g)
            sd[name + '.running_var'] = 0.5 + torch.rand(c, generator=g)
        else:
            sd[name + '.weight'], sd[name + '.bias'] = torch.ones(c), torch.zeros(c)
            sd[name + '.running_mean'], sd[name + '.running_var'] = torch.zeros(c), torch.ones(c)
        sd[name + '.num_batches_tracked'] = torch.zeros((), dtype=torch.long)

    def res(p, ci, co):
        if ci != co:
            conv(p + '.i_branch.0.weight', co, 1, ci)
        bn(p + '.conv_branch.0', ci)
        conv(p + '.conv_branch.2.weight', co, 3, ci)
        bn(p + '.conv_branch.3', co)
        conv(p + '.conv_branch.5.weight', co, 3, co)

    def ublock(p, planes):
        c = planes[0]
        for i in range(2):
            res(f'{p}.blocks.block{i}', c, c)
        if len(planes) > 1:
            bn(p + '.conv.0', c)
            conv(p + '.conv.2.weight', planes[1], 2, c)
            ublock(p + '.u', planes[1:])
            bn(p + '.deconv.0', planes[1])
            conv(p + '.deconv.2.weight', c, 2, planes[1])
            res(p + '.blocks_tail.block0', 2 * c, c)
            res(p + '.blocks_tail.block1', c, c)

    conv('input_conv.0.weight', channels, 3, dim_in)
    ublock('unet', [channels * (i + 1) for i in range(num_blocks)])
    bn('output_layer.0', channels)
    for head, co in (('semantic_linear', 2), ('offset_linear', 3)):
        b = (6.0 / (2 * channels)) ** 0.5
        sd[head + '.0.weight'] = (torch.rand((channels, channels), generator=g) * 2 - 1) * b
        sd[head + '.0.bias'] = 0.1 * torch.randn(channels, generator=g)
        bn(head + '.1', channels)
        sd[head + '.3.weight'] = 0.1 * torch.randn((co, channels), generator=g)
        sd[head + '.3.bias'] = 0.1 * torch.randn(co, generator=g)
    return sd
