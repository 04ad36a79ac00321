"""`TreeLearn` -- drop-in for `tree_learn.model.TreeLearn` (reference tree_learn/model/tree_learn.py:11-126).

Same constructor signature, `forward(batch, return_loss)` contract, output dict and `state_dict`
key/shape layout (b), so `load_checkpoint` of a reference `.pth` works and
tools/pipeline / tools/training can construct and call it unchanged.  What differs is what it
runs on: no spconv module tree -- parameters live in plain containers named like the reference's
modules, and the forward is a flat schedule of fused C-ABI kernels (treelearn_b200/csrc):

    point->voxel (Morton sort) -> level pyramid + rulebooks -> ~70 segmented gather-GEMM convs,
    each with residual add / skip-concat / 1x1 projection folded into the GEMM and the consumer's
    BatchNorm(eval)+ReLU folded into the epilogue -> voxel->point gather + both heads.

There is no CPU path: tensors are moved to the current CUDA device (like the reference's
`cuda_cast`, tree_learn/util/train.py:28-43) and the extension must be present.
"""
import functools
import os

import torch
import torch.nn.functional as F
from torch import nn

from . import _lib, sparse
from .sparse import Seg

LOSS_MULTIPLIER_SEMANTIC = 50     # reference tree_learn.py:9
BN_EPS, BN_MOMENTUM = 1e-4, 0.1   # reference tree_learn.py:34


class SparseConvWeight(nn.Module):
    """Parameter holder with spconv's KRSC layout [C_out, k, k, k, C_in]; bias-free."""

    def __init__(self, in_channels, out_channels, kernel_size):
        super().__init__()
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, kernel_size
        k = kernel_size
        self.weight = nn.Parameter(torch.empty(out_channels, k, k, k, in_channels))
        nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)


class HeadMLP(nn.Sequential):
    """Linear-BN-ReLU-Linear (reference blocks.py:8-26): keys 0.*, 1.*, 3.*."""

    def __init__(self, channels, out_channels, norm_fn):
        super().__init__(nn.Linear(channels, channels), norm_fn(channels), nn.ReLU(), nn.Linear(channels, out_channels))

    def init_weights(self):
        for m in self:
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                nn.init.constant_(m.bias, 0)
        nn.init.normal_(self[3].weight, 0, 0.01)
        nn.init.constant_(self[3].bias, 0)


def _container_path(root, path):
    """Walk/create plain nn.Module containers along a dotted path; returns (parent, leaf_name)."""
    parts = path.split('.')
    mod = root
    for p in parts[:-1]:
        if p not in mod._modules:
            mod.add_module(p, nn.Module())
        mod = mod._modules[p]
    return mod, parts[-1]


class TreeLearn(nn.Module):
    def __init__(self, channels=32, num_blocks=7, kernel_size=3, dim_coord=3, dim_feat=1, fixed_modules=[],
                 use_feats=True, use_coords=False, spatial_shape=None, max_num_points_per_voxel=3,
                 voxel_size=0.1, mode='fp32', **kwargs):
        super().__init__()
        if kernel_size != 3:
            raise NotImplementedError('treelearn_b200 builds the 3^3 submanifold rulebook only (reference default)')
        self.channels, self.num_blocks = channels, num_blocks
        self.dim_coord, self.dim_feat = dim_coord, dim_feat
        self.voxel_size = voxel_size
        self.fixed_modules = fixed_modules
        self.use_feats, self.use_coords = use_feats, use_coords
        self.spatial_shape = spatial_shape
        self.max_num_points_per_voxel = max_num_points_per_voxel
        # 'fp32' (SIMT, exact-ish) | 'tf32' | 'f16' (tcgen05, fp16 operands) | 'f16x2' (tcgen05, two-term fp16 split of both
        # operands = fp32-equivalent products: the reference's inference arithmetic is fp32)
        # 'mixed': f16x2 on the first `split_levels` U-Net levels (where 94 % of the voxels and of the output error live), f16
        # below; the activated tensors change format at that one level boundary (two small elementwise passes)
        self.mode = mode
        if mode not in ('fp32', 'tf32', 'f16', 'f16x2', 'mixed'):
            raise ValueError(f'unknown mode {mode!r}')
        self.split_levels = {'f16x2': num_blocks, 'mixed': int(kwargs.get('split_levels', os.environ.get('TL_SPLIT_LEVELS', '2')))}.get(mode, 0)
        if mode in ('f16', 'f16x2', 'mixed') and channels % 32 != 0:
            raise ValueError(f"mode={mode!r} needs channels % 32 == 0 (every conv must take the tcgen05 path)")
        self.planes = [channels * (i + 1) for i in range(num_blocks)]
        self._norm = functools.partial(nn.BatchNorm1d, eps=BN_EPS, momentum=BN_MOMENTUM)
        self._packed = None
        # measured on B200 (cfg2): building levels >= 1 on a side stream next to the level-0 convs gains nothing (17.10 vs 16.96 ms:
        # the persistent conv CTAs leave the small geometry kernels no room to run concurrently) -> off by default
        self.overlap_geometry = os.environ.get('TL_OVERLAP_GEOMETRY', '0') != '0'

        self._put('input_conv.0', SparseConvWeight(dim_coord + dim_feat, channels, 3))
        self._declare_ublock('unet', 0)
        self._put('output_layer.0', self._norm(channels))
        self.semantic_linear = HeadMLP(channels, 2, self._norm)
        self.offset_linear = HeadMLP(channels, 3, self._norm)
        self.init_weights()
        for name in fixed_modules:
            for p in getattr(self, name).parameters():
                p.requires_grad = False

    # ---- parameter tree with the reference's names ------------------------------------------------
    def _put(self, path, module):
        parent, leaf = _container_path(self, path)
        parent.add_module(leaf, module)

    def _get(self, path):
        mod = self
        for p in path.split('.'):
            mod = mod._modules[p]
        return mod

    def _declare_residual(self, p, c_in, c_out):
        if c_in != c_out:
            self._put(p + '.i_branch.0', SparseConvWeight(c_in, c_out, 1))
        self._put(p + '.conv_branch.0', self._norm(c_in))
        self._put(p + '.conv_branch.2', SparseConvWeight(c_in, c_out, 3))
        self._put(p + '.conv_branch.3', self._norm(c_out))
        self._put(p + '.conv_branch.5', SparseConvWeight(c_out, c_out, 3))

    def _declare_ublock(self, p, l):
        c = self.planes[l]
        for i in range(2):
            self._declare_residual(f'{p}.blocks.block{i}', c, c)
        if l + 1 < self.num_blocks:
            cn = self.planes[l + 1]
            self._put(p + '.conv.0', self._norm(c))
            self._put(p + '.conv.2', SparseConvWeight(c, cn, 2))
            self._declare_ublock(p + '.u', l + 1)
            self._put(p + '.deconv.0', self._norm(cn))
            self._put(p + '.deconv.2', SparseConvWeight(cn, c, 2))
            self._declare_residual(p + '.blocks_tail.block0', 2 * c, c)
            self._declare_residual(p + '.blocks_tail.block1', c, c)

    def init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.BatchNorm1d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
            elif isinstance(m, HeadMLP):
                m.init_weights()

    def train(self, mode=True):
        super().train(mode)
        for name in self.fixed_modules:      # BN of frozen modules stays in eval (reference tree_learn.py:66-72)
            for m in getattr(self, name).modules():
                if isinstance(m, nn.BatchNorm1d):
                    m.eval()
        return self

    # ---- weight packing (eval): BN folded to scale/shift, conv weights to the kernel layout -------
    def _version_key(self):
        return (self.mode, self.split_levels, self.input_conv._modules['0'].weight.device,
                tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers())))

    def _pack(self):
        key = self._version_key()
        if self._packed is not None and self._packed['key'] == key:
            return self._packed
        tf32 = self.mode in ('tf32', 'f16', 'mixed')
        half = self.mode in ('f16', 'mixed')
        half_modes = self.mode in ('f16', 'f16x2', 'mixed')

        def ts_of(name):   # (hi, lo) terms of the module's level: 2 = f16x2, 1 = f16, 0 = not a half mode
            if not half_modes:
                return 0
            if name.split('.').count('u') < self.split_levels:
                return 2
            return 1 if sparse.USE_TS else 0
        pk = {'key': key}
        with torch.no_grad():
            for name, m in self.named_modules():
                if isinstance(m, SparseConvWeight):
                    ts = ts_of(name)
                    w = m.weight.detach().float()
                    co, ci = m.out_channels, m.in_channels
                    w = w.reshape(co, -1, ci)
                    if name.endswith(('blocks_tail.block0.conv_branch.2', 'blocks_tail.block0.i_branch.0')):
                        # input is cat(identity, decoder) (reference blocks.py:146): one segment per half
                        pieces = [w[:, :, :ci // 2], w[:, :, ci // 2:]]
                    else:
                        pieces = [w]
                    out = []
                    for piece in pieces:
                        if ts and 'i_branch' in name and sparse.TS_KIND == 1:
                            # 1x1 projection of the fp32 residual stream (P-layout in and out): fp32 SIMT kernel in mode
                            # f16x2 (exact), round 1's TF32 kernel in mode f16; natural addressing, so both channel axes
                            # of the weight are permuted instead
                            wp = sparse.permute_p_weight(piece.permute(1, 0, 2))
                            out.append(wp.permute(0, 2, 1).contiguous() if ts == 2 else sparse.pack_weight_tc(wp, False))
                        elif ts and tc_eligible(piece.shape[2], co):
                            out.append(sparse.pack_weight(piece.permute(1, 0, 2), ts))      # f16 / f16x2 kernels: B slabs (+ lo terms)
                        elif half and tc_eligible(piece.shape[2], co) and 'i_branch' not in name:
                            out.append(sparse.pack_weight_tc(piece.permute(1, 0, 2), True))    # fp16 B-operand slabs
                        elif tf32 and tc_eligible(piece.shape[2], co):
                            out.append(sparse.pack_weight_tc(piece.permute(1, 0, 2), False))   # TF32 B-operand slabs
                        else:
                            out.append(piece.permute(1, 2, 0).contiguous())               # [K, Ci, Co] SIMT layout
                    pk[name] = out
                elif isinstance(m, nn.BatchNorm1d) and not name.startswith(('semantic_linear', 'offset_linear')):
                    s = (m.weight / torch.sqrt(m.running_var + m.eps)).float()
                    t = (m.bias - m.running_mean * s).float()
                    pk[name] = (s.contiguous(), t.contiguous())
            for head, tag in ((self.semantic_linear, 'sem'), (self.offset_linear, 'off')):
                bn = head[1]
                s = bn.weight / torch.sqrt(bn.running_var + bn.eps)
                t = bn.bias - bn.running_mean * s
                pk[tag + '_w1'] = (head[0].weight * s[:, None]).float().contiguous()
                pk[tag + '_b1'] = (head[0].bias * s + t).float().contiguous()
                pk[tag + '_w2'] = head[3].weight.float().contiguous()
                pk[tag + '_b2'] = head[3].bias.float().contiguous()
        self._packed = pk
        return pk

    # ---- forward ------------------------------------------------------------------------------------
    def forward(self, batch, return_loss):
        voxel_out, v2p = self.forward_backbone(**batch)
        output = self.forward_head(voxel_out, v2p)
        if return_loss:
            output = self.get_loss(model_output=output, **batch)
        return output

    def forward_backbone(self, coords, input_feats, batch_ids, batch_size, **kwargs):
        dev = torch.device('cuda', torch.cuda.current_device())
        if int(batch_size) == 1 and batch_ids.device.type == 'cpu':
            # one tile per batch (the pipeline's setting, configs/pipeline/pipeline.yaml:15-17): every id is 0 by construction
            # of collate_fn (dataset.py:214-226) -- fill on the device instead of copying 8 B per point over PCIe
            batch_ids = torch.zeros(batch_ids.shape[0], dtype=torch.int64, device=dev)
        coords, input_feats, batch_ids = (t.to(dev, non_blocking=True) for t in (coords, input_feats, batch_ids))
        vfeats, vcoords, keys, v2p = sparse.voxelize(
            coords, input_feats, batch_ids, batch_size, self.voxel_size, self.use_coords, self.use_feats,
            self.max_num_points_per_voxel)
        if self.spatial_shape is not None:
            shape = [int(s) for s in self.spatial_shape]
        else:
            shape = (vcoords[:, 1:].max(dim=0).values + 1).tolist()   # reference tree_learn.py:165
        if self._needs_autograd_path():
            levels = sparse.build_levels(keys, vcoords, shape, self.num_blocks)
            return self._train_backbone(vfeats, levels), v2p
        # inference: levels >= 1 are built on a side stream while the level-0 convolutions already run
        levels = sparse.LazyLevels(keys, vcoords, shape, self.num_blocks) if self.overlap_geometry \
            else sparse.build_levels(keys, vcoords, shape, self.num_blocks)
        return self._run_backbone(vfeats, levels), v2p

    def _needs_autograd_path(self):
        """The fused inference schedule folds BN running statistics into the conv epilogues and records no graph:
        it is only valid when no BatchNorm uses batch statistics and no gradient is wanted."""
        if any(m.training for m in self.modules() if isinstance(m, nn.BatchNorm1d)):
            return True
        return torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())

    # ---- training / autograd schedule (unfused: BN batch statistics sit between the convs) ----------
    def _train_backbone(self, vfeats, levels):
        from . import autograd as ag
        mode = _lib.MODE_FP32 if self.mode == 'fp32' else _lib.MODE_TF32
        geoms = {'subm': [ag.subm_geom(lv) for lv in levels],
                 'down': [ag.down_geom(levels[l], levels[l + 1]) for l in range(len(levels) - 1)],
                 'up': [ag.up_geom(levels[l], levels[l + 1]) for l in range(len(levels) - 1)]}
        x = ag.sparse_conv(vfeats, self._get('input_conv.0').weight, geoms['subm'][0], mode)
        x = self._train_ublock('unet', 0, x, geoms, mode)
        return ag.bn_relu(x, self._get('output_layer.0'))

    def _train_residual(self, p, x, geom, mode):
        from . import autograd as ag

        def conv(a, weight, g):
            """Inputs wider than the conv's output (the 2C skip concat of blocks_tail.block0) run as two convs over the
            channel halves: the kernels take at most 256 channels per operand, and the halves are what the fused inference
            schedule feeds as two segments anyway."""
            ci, co = weight.shape[-1], weight.shape[0]
            if ci == 2 * co:
                return (ag.sparse_conv(a[:, :co].contiguous(), weight[..., :co], g, mode) +
                        ag.sparse_conv(a[:, co:].contiguous(), weight[..., co:], g, mode))
            return ag.sparse_conv(a, weight, g, mode)

        h = conv(ag.bn_relu(x, self._get(p + '.conv_branch.0')), self._get(p + '.conv_branch.2').weight, geom)
        h = conv(ag.bn_relu(h, self._get(p + '.conv_branch.3')), self._get(p + '.conv_branch.5').weight, geom)
        blk = self._get(p)
        if 'i_branch' in blk._modules:   # 1x1 projection of the residual branch (reference blocks.py:29-39)
            x = conv(x, self._get(p + '.i_branch.0').weight, ag.identity_geom(x.shape[0]))
        return h + x

    def _train_ublock(self, p, l, x, geoms, mode):
        from . import autograd as ag
        g = geoms['subm'][l]
        for i in range(2):
            x = self._train_residual(f'{p}.blocks.block{i}', x, g, mode)
        if l + 1 < self.num_blocks:
            d = ag.sparse_conv(ag.bn_relu(x, self._get(p + '.conv.0')), self._get(p + '.conv.2').weight, geoms['down'][l], mode)
            d = self._train_ublock(p + '.u', l + 1, d, geoms, mode)
            up = ag.sparse_conv(ag.bn_relu(d, self._get(p + '.deconv.0')), self._get(p + '.deconv.2').weight, geoms['up'][l], mode)
            x = torch.cat([x, up], dim=1)
            for i in range(2):
                x = self._train_residual(f'{p}.blocks_tail.block{i}', x, g, mode)
        return x

    def _run_backbone(self, vfeats, levels):
        pk = self._pack()
        mode = self._mode_id(0)
        g0 = levels[0]
        x, xa = sparse.conv([Seg(vfeats, pk['input_conv.0'][0], g0.nbr, g0.nbr_mask)], g0.n, self.planes[0], mode,
                            raw=True, act1=pk['unet.blocks.block0.conv_branch.0'])
        return self._run_ublock('unet', 0, x, xa, levels, pk, pk['output_layer.0'])

    def _mode_id(self, l):
        """Kernel mode of U-Net level l."""
        if self.mode in ('f16', 'f16x2', 'mixed'):
            return _lib.MODE_F16X2 if l < self.split_levels else _lib.MODE_F16
        return {'fp32': _lib.MODE_FP32, 'tf32': _lib.MODE_TF32}[self.mode]

    def _run_ublock(self, p, l, x, xa, levels, pk, ret_act):
        """x: raw features [n_l, C_l]; xa = relu(bn(x)) for blocks.block0; returns relu(ret_bn(level output))."""
        mode = self._mode_id(l)
        g, c, n = levels[l], self.planes[l], levels[l].n
        nbr = lambda src, w: Seg(src, w, g.nbr, g.nbr_mask)   # noqa: E731
        conv = functools.partial(sparse.conv, n_out=n, c_out=c, mode=mode)
        b0, b1 = p + '.blocks.block0.conv_branch', p + '.blocks.block1.conv_branch'
        ha = conv([nbr(xa, pk[b0 + '.2'][0])], act1=pk[b0 + '.3'])
        y, ya = conv([nbr(ha, pk[b0 + '.5'][0])], residual=x, raw=True, act1=pk[b1 + '.0'])
        ha = conv([nbr(ya, pk[b1 + '.2'][0])], act1=pk[b1 + '.3'])
        if l + 1 == self.num_blocks:
            return conv([nbr(ha, pk[b1 + '.5'][0])], residual=y, act1=ret_act)
        t0, t1 = p + '.blocks_tail.block0', p + '.blocks_tail.block1.conv_branch'
        s_cat, t_cat = pk[t0 + '.conv_branch.0']              # BN over the 2C concat: halves go to their producers
        z, za_down, za_tail = conv([nbr(ha, pk[b1 + '.5'][0])], residual=y, raw=True, act1=pk[p + '.conv.0'],
                                   act2=(s_cat[:c], t_cat[:c]))
        gn, cn = levels[l + 1], self.planes[l + 1]
        d, da = sparse.conv([Seg(za_down, pk[p + '.conv.2'][0], g.down_index, g.down_mask)], gn.n, cn, mode,
                            raw=True, act1=pk[p + '.u.blocks.block0.conv_branch.0'])
        child = self._mode_id(l + 1)
        if child != mode:      # f16x2 -> f16 at this level boundary: hi + lo -> one fp16 term (the fp32 stream `d` is format-free)
            da = sparse.split_to_half(da)
        ua = self._run_ublock(p + '.u', l + 1, d, da, levels, pk, pk[p + '.deconv.0'])
        if child != mode:      # the child's fp16 output as a (hi, lo = 0) operand of this level's inverse conv
            ua = sparse.half_to_split(ua)
        e, ea = conv([Seg(ua, pk[p + '.deconv.2'][0], g.up_index, g.up_mask)], raw=True, act1=(s_cat[c:], t_cat[c:]))
        wa, wi = pk[t0 + '.conv_branch.2'], pk[t0 + '.i_branch.0']
        ha = conv([nbr(za_tail, wa[0]), nbr(ea, wa[1])], act1=pk[t0 + '.conv_branch.3'])
        if mode in (_lib.MODE_F16, _lib.MODE_F16X2) and sparse.TS_KIND != 2:
            # the 1x1 projection reads the fp32 residual-stream tensors: run it as its own launch (TF32 tensor cores in mode
            # f16, fp32 FMA in mode f16x2) and feed it in as the residual of the fp16-operand 3^3 conv
            pmode = _lib.MODE_FP32 if (mode == _lib.MODE_F16X2) else _lib.MODE_TF32
            proj = sparse.conv([Seg(z, wi[0]), Seg(e, wi[1])], n, c, pmode, raw=True)
            t, ta = conv([nbr(ha, pk[t0 + '.conv_branch.5'][0])], residual=proj, raw=True, act1=pk[t1 + '.0'])
        else:
            t, ta = conv([nbr(ha, pk[t0 + '.conv_branch.5'][0]), Seg(z, wi[0]), Seg(e, wi[1])], raw=True,
                         act1=pk[t1 + '.0'])
        ha = conv([nbr(ta, pk[t1 + '.2'][0])], act1=pk[t1 + '.3'])
        return conv([nbr(ha, pk[t1 + '.5'][0])], residual=t, act1=ret_act)

    def forward_head(self, voxel_out, v2p):
        if voxel_out.requires_grad or self.semantic_linear[1].training or self.offset_linear[1].training:
            # training: gather + the two tiny MLPs as torch ops so autograd and batch-stat BN apply (tree_learn.py:97-103)
            feats = voxel_out.float()[v2p]
            return {'backbone_feats': feats, 'semantic_prediction_logits': self.semantic_linear(feats),
                    'offset_predictions': self.offset_linear(feats)}
        feats, logits, offs = sparse.heads(voxel_out, v2p, self._pack(), split=self._mode_id(0) == _lib.MODE_F16X2)
        return {'backbone_feats': feats, 'semantic_prediction_logits': logits, 'offset_predictions': offs}

    def get_loss(self, model_output, semantic_labels, offset_labels, masks_off, masks_sem, **kwargs):
        dev = model_output['offset_predictions'].device
        logits = model_output['semantic_prediction_logits'].float()
        offs = model_output['offset_predictions'].float()
        sem, off = point_wise_loss(logits, offs, masks_sem.to(dev), masks_off.to(dev), semantic_labels.to(dev),
                                   offset_labels.to(dev))
        loss_dict = {'semantic_loss': sem * LOSS_MULTIPLIER_SEMANTIC, 'offset_loss': off}
        return sum(loss_dict.values()), loss_dict


def point_wise_loss(semantic_prediction_logits, offset_predictions, masks_sem, masks_off, semantic_labels,
                    offset_labels, weights=None):
    """Masked CE (sum / count) and mean L2 offset error -- reference tree_learn/util/train.py:145-166 (stays torch)."""
    dev = semantic_prediction_logits.device
    masks_sem, masks_off = masks_sem.to(dev), masks_off.to(dev)
    semantic_labels, offset_labels = semantic_labels.to(dev), offset_labels.to(dev)
    n_sem = int(masks_sem.sum())
    if n_sem == 0:
        semantic_loss = 0 * semantic_prediction_logits.sum()
    else:
        ce = F.cross_entropy(semantic_prediction_logits[masks_sem], semantic_labels[masks_sem],
                             reduction='sum' if weights is None else 'none')
        semantic_loss = (ce if weights is None else (ce * weights).sum()) / n_sem
    if int(masks_off.sum()) == 0:
        offset_loss = 0 * offset_predictions.sum()
    else:
        diff = offset_predictions[masks_off] - offset_labels[masks_off]
        offset_loss = diff.pow(2).sum(1).sqrt().mean()
    return semantic_loss, offset_loss


def tc_eligible(c_in, c_out):
    """Same rule as csrc/tl_conv_tc.cu: the tcgen05 path takes 32-channel K blocks and N = C_out <= 256."""
    return c_in % 32 == 0 and c_in <= 256 and c_out % 32 == 0 and c_out <= 256


def _round_tf32(w):
    """Round-to-nearest-even to TF32 (10 explicit mantissa bits) so the tensor core's operand truncation is exact."""
    i = w.contiguous().view(torch.int32)
    r = ((i + 0x0FFF + ((i >> 13) & 1)) & ~0x1FFF)
    return r.view(torch.float32)
