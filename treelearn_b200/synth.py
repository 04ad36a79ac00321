"""Synthetic forest tiles (SURVEY.md §8d generator spec) -- the bench/test workload.

There is no network for the L1W data, so BASELINE.json's configs are quoted on these tiles:
ground sheet + trunks (cylinder surfaces) + noisy branches, down-sampled to one point per 0.1 m
voxel (what `generate_tiles` does with open3d, tree_learn/util/pipeline.py:43-46), rounded to
2 decimals and xy-centred.  Output mirrors `TreeDataset.collate_fn`
(tree_learn/dataset/dataset.py:214-226).
"""
import numpy as np
import torch

TREE_CLASS, NON_TREE_CLASS = 0, 1   # tree_learn/dataset/dataset.py:9-10


def synth_forest(edge=20.0, height=20.0, n_trees=20, seed=0, voxel=0.1, ground_density=1000.0):
    """Returns dict(coords f32 [N,3] centred, feat f32 [N] verticality, inst i64 [N] (0 = ground),
    base f32 [N,3] tree-base position per point, n_raw)."""
    rng = np.random.default_rng(seed)
    pts, inst, vert, base = [], [], [], []

    def ground_z(x, y):
        return 1.0 + 0.3 * np.sin(x / 3.0) + 0.3 * np.cos(y / 4.0)

    ng = int(ground_density * edge * edge)
    gx, gy = rng.uniform(0, edge, ng), rng.uniform(0, edge, ng)
    pts.append(np.stack([gx, gy, ground_z(gx, gy) + rng.normal(0, 0.03, ng)], 1))
    inst.append(np.zeros(ng, np.int64))
    vert.append(rng.uniform(0, 1, ng))
    base.append(np.zeros((ng, 3)))
    for t in range(n_trees):
        cx, cy = rng.uniform(1, edge - 1, 2)
        h = rng.uniform(10, 18)
        r = rng.uniform(0.1, 0.3)
        b = np.array([cx, cy, ground_z(cx, cy)])
        nt = int(400 * h)
        ang = rng.uniform(0, 2 * np.pi, nt)
        tz = rng.uniform(1.0, 0.6 * h, nt)
        pts.append(np.stack([cx + r * np.cos(ang), cy + r * np.sin(ang), tz], 1))
        inst.append(np.full(nt, t + 1, np.int64))
        vert.append(rng.uniform(0.7, 1.0, nt))
        base.append(np.broadcast_to(b, (nt, 3)))
        nb, npb = 40, 300
        d = rng.normal(size=(nb, 3))
        d[:, 2] = np.abs(d[:, 2]) * 0.4
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        length = rng.uniform(1.0, 3.5, nb)
        z0 = rng.uniform(0.4 * h, h, nb)
        s = rng.uniform(0, 1, (nb, npb)) * length[:, None]
        p = np.stack([cx + s * d[:, None, 0], cy + s * d[:, None, 1], z0[:, None] + s * d[:, None, 2]], 2)
        p = p + rng.normal(size=p.shape) * (0.08 * (0.3 + s / length[:, None]))[:, :, None]
        p = p.reshape(-1, 3)
        pts.append(p)
        inst.append(np.full(len(p), t + 1, np.int64))
        vert.append(rng.uniform(0, 1, len(p)))
        base.append(np.broadcast_to(b, (len(p), 3)))
    pts, inst, vert, base = (np.concatenate(a) for a in (pts, inst, vert, base))
    keep = np.all((pts >= 0) & (pts < np.array([edge, edge, height])), axis=1)
    pts, inst, vert, base = pts[keep], inst[keep], vert[keep], base[keep]
    n_raw = len(pts)
    c = np.floor(pts / voxel).astype(np.int64)
    key = (c[:, 0] << 40) | (c[:, 1] << 20) | c[:, 2]
    _, first = np.unique(key, return_index=True)
    first.sort()
    pts, inst, vert, base = pts[first], inst[first], vert[first], base[first]
    pts = np.round(pts, 2)
    centre = np.array([edge / 2, edge / 2, 0.0])
    return dict(coords=(pts - centre).astype(np.float32), feat=vert.astype(np.float32), inst=inst,
                base=(base - centre).astype(np.float32), n_raw=n_raw, centre=centre.astype(np.float32))


def make_batch(tiles, inner_edge=8.0):
    """Collate tiles (list of synth_forest dicts) into the model's input dict."""
    coords, feats, bids, sem, inst, off, mi, mo, ms, cen = [], [], [], [], [], [], [], [], [], []
    for b, t in enumerate(tiles):
        n = len(t['coords'])
        xyz = torch.from_numpy(t['coords'])
        tree = torch.from_numpy(t['inst'] > 0)
        inner = (xyz[:, 0].abs() < inner_edge / 2) & (xyz[:, 1].abs() < inner_edge / 2)
        o = torch.from_numpy(t['base']) - xyz
        o[~tree] = 0
        coords.append(xyz)
        feats.append(torch.from_numpy(t['feat']).reshape(-1, 1))
        bids.append(torch.full((n,), b, dtype=torch.long))
        sem.append(torch.where(tree, TREE_CLASS, NON_TREE_CLASS).long())
        inst.append(torch.from_numpy(t['inst']).long())
        off.append(o.float())
        mi.append(inner)
        ms.append(inner.clone())
        mo.append(inner & tree)
        cen.append(torch.from_numpy(t['centre']).reshape(1, 3).expand(n, 3))
    return {'coords': torch.cat(coords).float(), 'input_feats': torch.cat(feats).float(),
            'batch_ids': torch.cat(bids), 'semantic_labels': torch.cat(sem), 'instance_labels': torch.cat(inst),
            'masks_inner': torch.cat(mi), 'masks_off': torch.cat(mo), 'masks_sem': torch.cat(ms),
            'offset_labels': torch.cat(off), 'batch_size': len(tiles), 'centers': torch.cat(cen).float()}


def plot_tiles(n_side=8, inner_edge=8.0, outer_edge=13.5, stride=0.5, seed=7, trees_per_100m2=5.6, ground_density=1000.0):
    """BASELINE.json config 4: ONE synthetic plot cut into n_side x n_side overlapping tiles (inner squares of `inner_edge`
    stepped by stride * inner_edge, `outer_edge` of context on every side => 35 m tiles at the defaults, every interior
    point in (1 / stride)^2 = 4 inner squares), as `SampleGenerator.tile_generate_and_save` lays them out
    (tree_learn/util/data_preparation.py:364-431).  Returns the list of model input dicts (TreeDataset test-mode fields the
    inference path reads: coords centred on the tile, input_feats, batch_ids, batch_size, masks_inner, centers)."""
    step = stride * inner_edge
    span = (n_side - 1) * step + inner_edge                    # union of the inner squares
    edge = span + 2 * outer_edge
    f = synth_forest(edge=edge, n_trees=int(round(trees_per_100m2 * edge * edge / 100.0)), seed=seed, ground_density=ground_density)
    xyz, feat = f['coords'], f['feat']                         # plot coordinates, xy-centred
    first = -span / 2 + inner_edge / 2
    tiles = []
    for iy in range(n_side):
        for ix in range(n_side):
            cx, cy = np.float32(first + ix * step), np.float32(first + iy * step)
            half = np.float32(inner_edge / 2 + outer_edge)
            sel = (np.abs(xyz[:, 0] - cx) <= half) & (np.abs(xyz[:, 1] - cy) <= half)
            p = xyz[sel] - np.array([cx, cy, 0], dtype=np.float32)
            n = len(p)
            inner = (np.abs(p[:, 0]) <= inner_edge / 2) & (np.abs(p[:, 1]) <= inner_edge / 2)
            tree = f['inst'][sel] > 0
            off = (f['base'][sel] - xyz[sel]).astype(np.float32)
            off[~tree] = 0
            tiles.append({'coords': torch.from_numpy(p), 'input_feats': torch.from_numpy(feat[sel]).reshape(-1, 1),
                          'batch_ids': torch.zeros(n, dtype=torch.long), 'batch_size': 1,
                          'semantic_labels': torch.from_numpy(np.where(tree, TREE_CLASS, NON_TREE_CLASS)).long(),
                          'offset_labels': torch.from_numpy(off),
                          'masks_inner': torch.from_numpy(inner),
                          'centers': torch.from_numpy(np.array([cx, cy, 0], dtype=np.float32)).reshape(1, 3).expand(n, 3).contiguous()})
    return tiles


# named workloads (BASELINE.json configs; sizes per SURVEY §8d)
WORKLOADS = {
    'cfg1_200k': dict(edge=20.0, n_trees=20, seed=0),
    'cfg2_2M': dict(edge=60.0, n_trees=225, seed=1),
    'tiny': dict(edge=6.0, n_trees=2, seed=3, ground_density=300.0),
    'small': dict(edge=10.0, n_trees=5, seed=4),
}


def workload(name):
    return synth_forest(**WORKLOADS[name])


def randomize_bn_stats(model, seed=0):
    """Random-init weights keep BN as identity; randomise affine + running stats so BN is not a no-op."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                c = m.num_features
                m.weight.copy_(0.5 + torch.rand(c, generator=g))
                m.bias.copy_(0.2 * torch.randn(c, generator=g))
                m.running_mean.copy_(0.1 * torch.randn(c, generator=g))
                m.running_var.copy_(0.5 + torch.rand(c, generator=g))
    return model


def fit_probe_heads(model, batch, ridge=1e-3):
    """"Trained-like" heads for a random-init backbone (there is no network for the reference's checkpoint): the last
    Linear of both heads is fitted by ridge regression on THIS model's own hidden activations so that
      * offset_predictions approximate the tile's offset labels on tree points (vectors of several metres, like a trained
        model's: tree_learn/dataset/dataset.py:111-140), and
      * the semantic logits separate tree / non-tree points.
    Everything before the last Linear stays as initialised, so the backbone arithmetic that parity tests probe is untouched.
    `model` must be on the GPU in eval mode; the fitted weights are written into `model` and also returned as a dict of
    CPU tensors keyed like the state_dict ('offset_linear.3.weight', ...)."""
    dev = next(model.parameters()).device
    with torch.no_grad():
        out = model(batch, return_loss=False)
        feats = out['backbone_feats'].double()
        fitted = {}
        tree = (batch['semantic_labels'] == TREE_CLASS).to(dev)
        for name, head, target, rows in (
                ('offset_linear', model.offset_linear, batch['offset_labels'].to(dev).double(), tree),
                ('semantic_linear', model.semantic_linear,
                 torch.where(tree, 4.0, -4.0)[:, None].double() * torch.tensor([1.0, -1.0], device=dev, dtype=torch.float64),
                 torch.ones_like(tree))):
            bn = head[1]
            h = head[0].weight.double() @ feats.T + head[0].bias.double()[:, None]
            h = (h - bn.running_mean.double()[:, None]) / torch.sqrt(bn.running_var.double()[:, None] + bn.eps)
            h = torch.relu(h * bn.weight.double()[:, None] + bn.bias.double()[:, None]).T            # [N, C]
            a = torch.cat([h[rows], torch.ones((int(rows.sum()), 1), device=dev, dtype=torch.float64)], 1)
            gram = a.T @ a + ridge * len(a) * torch.eye(a.shape[1], device=dev, dtype=torch.float64)
            sol = torch.linalg.solve(gram, a.T @ target[rows])                                       # [C + 1, out]
            w, b = sol[:-1].T.float().contiguous(), sol[-1].float().contiguous()
            head[3].weight.copy_(w)
            head[3].bias.copy_(b)
            fitted[name + '.3.weight'], fitted[name + '.3.bias'] = w.cpu(), b.cpu()
    return fitted


class TrainedLikeOutputs(torch.nn.Module):
    """Stand-in for the trained part of the network that cannot be downloaded here: wraps a model and adds, per prepared
    batch, the fixed correction   labels - keep * (the model's own outputs at prepare time)   to its offsets and logits, so
    that what the clustering stage sees is  labels + (1 - keep) * model output  -- predictions `keep` of the way to the
    labels, with the model's metre-scale output as the residual error -- and the offset-shifted clustering / kNN stages do
    the work they do behind a trained network (tens of trees per tile, a large remaining-point set).  The wrapped model
    still runs in full every call; the correction is two element-wise adds.  Used by bench.py and the trained-scale tests;
    the comparison CUDA-vs-oracle is unaffected because both sides get the same correction."""

    def __init__(self, model, keep=0.9):
        super().__init__()
        self.model, self.keep, self._corr = model, keep, {}

    def prepare(self, batch):
        """Compute and keep the correction for `batch` (looked up by the identity of its coords tensor)."""
        with torch.no_grad():
            out = self.model(batch, return_loss=False)
            dev = out['offset_predictions'].device
            tree = (batch['semantic_labels'] == TREE_CLASS).to(dev)
            c_off = torch.where(tree[:, None], batch['offset_labels'].to(dev) - self.keep * out['offset_predictions'].float(),
                                torch.zeros(1, device=dev))
            want = torch.where(tree, 4.0, -4.0)[:, None] * torch.tensor([1.0, -1.0], device=dev)
            c_sem = want - self.keep * out['semantic_prediction_logits'].float()
        self._corr[id(batch['coords'])] = (c_off.contiguous(), c_sem.contiguous())

    def share(self, batch, like):
        """`batch` holds the same tile as the prepared batch `like` (e.g. its device-resident copy)."""
        self._corr[id(batch['coords'])] = self._corr[id(like['coords'])]

    def eval(self):
        self.model.eval()
        return self

    def forward(self, batch, return_loss=False):
        out = dict(self.model(batch, return_loss=return_loss))
        key = id(batch['coords'])
        src = key if key in self._corr else batch.get('_source_coords_id')
        if isinstance(src, (list, tuple)):          # several prepared tiles collated into one batch
            c_off = torch.cat([self._corr[k][0] for k in src])
            c_sem = torch.cat([self._corr[k][1] for k in src])
        else:
            c_off, c_sem = self._corr[src]
        out['offset_predictions'] = out['offset_predictions'] + c_off
        out['semantic_prediction_logits'] = out['semantic_prediction_logits'] + c_sem
        return out
