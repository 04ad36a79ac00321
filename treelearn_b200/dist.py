"""Multi-GPU plumbing for the two places the path shards.  One process per GPU, torch.distributed.

Inference: tiles are independent units (tree_learn/util/pipeline.py:83-103) => tile i goes to one rank (longest-
processing-time-first by point count), every rank crops its tiles to the inner square on the device, ONE variable-size
all-gather exchanges the inner rows, and the overlap merge (`ensemble`) + clustering run replicated on the merged
plot (the reference clusters the whole plot at once, tools/pipeline/pipeline.py:89-94).
The reference has no distributed code at all; these helpers are new.  They work with the nccl (GPU) and gloo (CPU
tests of the host logic) backends.
"""
import torch
import torch.distributed as dist


def shard_indices(weights, rank, world):
    """Greedy longest-processing-time-first partition of items (tiles) by weight; returns this rank's item indices
    in ascending order.  Deterministic: every rank computes the same assignment."""
    order = sorted(range(len(weights)), key=lambda i: (-float(weights[i]), i))
    load = [0.0] * world
    mine = []
    for i in order:
        r = min(range(world), key=lambda j: (load[j], j))
        load[r] += float(weights[i])
        if r == rank:
            mine.append(i)
    return sorted(mine)


def allgather_rows(t, group=None):
    """Variable-length all-gather along dim 0: every rank passes [n_r, ...] and gets cat over ranks (rank order)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return t
    world = dist.get_world_size(group)
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes)
    pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)


def allgather_rows_known(t, sizes, group=None):
    """Variable-length all-gather along dim 0 when every rank already knows all row counts `sizes` (one per rank):
    ONE collective (`all_gather_into_tensor`) on a buffer sized for the largest contribution, no size exchange."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return t
    world = dist.get_world_size(group)
    assert len(sizes) == world and t.shape[0] == sizes[dist.get_rank(group)]
    m = max(max(sizes), 1)
    pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    out = torch.empty((world * m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return torch.cat([out[r * m:r * m + s] for r, s in enumerate(sizes)], dim=0)


def _forward_chunk(model, chunk, dev):
    """One network forward over several tiles as ONE batch (tile = batch element, like a DataLoader batch of the reference's
    collate_fn, dataset.py:214-226) and the inner-square rows [x, y, z, logits(2), offsets(3), verticality] of all of them.
    A 35 m tile of ~0.7 M points spends about half of its forward in per-launch floors of the deep U-Net levels (~40 us x
    66 conv launches) and host synchronisations of the level builder; several tiles per forward share them."""
    coords = torch.cat([b['coords'].to(dev, non_blocking=True) for b in chunk])
    feats = torch.cat([b['input_feats'].to(dev, non_blocking=True) for b in chunk])
    counts = [int(b['coords'].shape[0]) for b in chunk]
    ids = torch.repeat_interleave(torch.arange(len(chunk), device=dev), torch.tensor(counts, device=dev), output_size=sum(counts))
    batch = {'coords': coords, 'input_feats': feats, 'batch_ids': ids, 'batch_size': len(chunk),
             '_source_coords_id': [id(b['coords']) for b in chunk]}          # unknown keys are ignored by the model
    out = model(batch, return_loss=False)
    inner = torch.cat([b['masks_inner'] for b in chunk]).to(dev, non_blocking=True)
    centers = torch.cat([b['centers'].to(dev, non_blocking=True) for b in chunk])
    xyz = (coords + centers)[inner]
    return torch.cat([xyz, out['semantic_prediction_logits'][inner], out['offset_predictions'][inner], feats[inner][:, -1:]], dim=1)


def segment_plot(model, tiles, grouping_cfg, group=None, marks=None, points_per_forward=6_000_000):
    """Whole-plot inference (BASELINE.json config 4): `tiles` is the same list of host batch dicts on every rank.
    Returns (merged coords [P,3], instance labels [P], n_clusters) as CUDA tensors, identical on every rank.
    `marks` (optional list) receives (name, cuda event) pairs at the stage boundaries: forward, allgather, merge, cluster.
    A rank runs its tiles in chunks of up to `points_per_forward` points per network forward (at least one tile)."""
    from . import pipeline
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    owner = [None] * len(tiles)
    for r in range(world):
        for i in shard_indices([t['coords'].shape[0] for t in tiles], r, world):
            owner[i] = r
    mine = [i for i, r in enumerate(owner) if r == rank]
    # every rank holds every tile's inner mask, so all row counts are known without a size exchange
    sizes = [sum(int(tiles[i]['masks_inner'].sum()) for i in range(len(tiles)) if owner[i] == r) for r in range(world)]

    def mark(name):
        if marks is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            marks.append((name, e))

    rows = []
    dev = torch.device('cuda', torch.cuda.current_device())
    mark('start')
    with torch.no_grad():
        model.eval()
        chunk, pts = [], 0
        for i in mine + [None]:
            n = 0 if i is None else int(tiles[i]['coords'].shape[0])
            if chunk and (i is None or pts + n > points_per_forward):
                rows.append(_forward_chunk(model, chunk, dev))
                chunk, pts = [], 0
            if i is not None:
                chunk.append(tiles[i])
                pts += n
    local = torch.cat(rows) if rows else torch.zeros((0, 9), device=dev)
    mark('forward')
    allrows = allgather_rows_known(local.contiguous(), sizes, group)          # the one collective of the inference path
    mark('allgather')
    coords, vals = pipeline.ensemble_cuda(allrows[:, :3].contiguous(), allrows[:, 3:].contiguous())
    mark('merge')
    labels, n_clusters = pipeline.instances_cuda(coords, vals[:, 2:5].contiguous(), vals[:, 0:2].contiguous(),
                                                 vals[:, 5].contiguous(), grouping_cfg)
    mark('cluster')
    return coords, labels, n_clusters


# ---- data-parallel training (BASELINE.json config 5) -----------------------------------------------
def grad_buckets(model):
    """Parameters grouped for the gradient all-reduce: one bucket per U-Net level (depth = number of '.u.' hops in
    the reference's module path, tree_learn/model/blocks.py:97-135) plus one for the input conv / output layer /
    heads.  Ordered deepest level first -- the order in which backward finishes their weight gradients last-to-first
    is irrelevant for correctness; the split keeps every NCCL call in the tens-of-MB range (level 6 alone holds
    ~11 M of the 30 M parameters)."""
    buckets = {}
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        depth = name.split('.').count('u') if name.startswith('unet.') else -1
        buckets.setdefault(depth, []).append(p)
    return [buckets[d] for d in sorted(buckets, reverse=True)]


def allreduce_grads(model, group=None):
    """Average `.grad` over ranks (identical replicas, per-GPU BatchNorm statistics -- the reference has no SyncBN):
    one flat all-reduce(sum) per bucket, launched asynchronously and waited together, then scaled by 1/world."""
    if not (dist.is_available() and dist.is_initialized()):
        return
    world = dist.get_world_size(group)
    if world == 1:
        return
    pending = []
    for params in grad_buckets(model):
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in params]
        flat = torch.cat([g.reshape(-1) for g in grads])
        pending.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True), flat, params))
    for work, flat, params in pending:
        work.wait()
        flat.mul_(1.0 / world)
        off = 0
        for p in params:
            n = p.numel()
            p.grad = flat[off:off + n].view_as(p).clone() if p.grad is None else p.grad.copy_(flat[off:off + n].view_as(p))
            off += n


class OverlappedGradReducer:
    """Gradient all-reduce launched DURING backward (SURVEY §8e training row): every parameter carries a
    post-accumulate-grad hook; when the last gradient of a bucket (one U-Net level, `grad_buckets`) has been written the
    bucket is flattened and its all-reduce(sum) starts asynchronously on the communicator's stream while autograd keeps
    running the shallower levels' data- and weight-gradient kernels.  Backward reaches the deepest level first on the way
    down the decoder side and finishes level 0 last, so only the last bucket's transfer is exposed.
    `finish()` launches whatever never completed (parameters without a gradient count as zeros), waits, scales by 1/world
    and writes the averages back into `.grad`.  With world == 1 (or no process group) it does nothing."""

    def __init__(self, model, group=None):
        self.group = group
        self.active = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.world = dist.get_world_size(group) if self.active else 1
        self.buckets = grad_buckets(model)
        self._where = {}
        self._handles = []
        for b, params in enumerate(self.buckets):
            for p in params:
                self._where[id(p)] = b
                if self.active:
                    self._handles.append(p.register_post_accumulate_grad_hook(self._hook))
        self.reset()

    def reset(self):
        self._ready = [0] * len(self.buckets)
        self._pending = [None] * len(self.buckets)
        self.launched_in_backward = 0

    def _launch(self, b):
        params = self.buckets[b]
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])
        self._pending[b] = (dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True), flat)

    def _hook(self, p):
        b = self._where[id(p)]
        self._ready[b] += 1
        if self._ready[b] == len(self.buckets[b]) and self._pending[b] is None:
            self._launch(b)
            self.launched_in_backward += 1

    def finish(self):
        if not self.active:
            return
        for b in range(len(self.buckets)):
            if self._pending[b] is None:
                self._launch(b)
        for b, params in enumerate(self.buckets):
            work, flat = self._pending[b]
            work.wait()
            flat.mul_(1.0 / self.world)
            off = 0
            for p in params:
                n = p.numel()
                g = flat[off:off + n].view_as(p)
                if p.grad is None:
                    p.grad = g.clone()
                else:
                    p.grad.copy_(g)
                off += n
        self.reset()

    def remove(self):
        for h in self._handles:
            h.remove()
        self._handles = []


def train_step(model, optimizer, batch, group=None, grad_clip=None, reducer=None, scaler=None, autocast=False):
    """One data-parallel training step on this rank's batch (tools/training/train.py:32-44): forward + loss (optionally
    under fp16 autocast with a GradScaler, as the reference trains), backward with the bucketed gradient all-reduce
    overlapped (`reducer`, an OverlappedGradReducer; None = all-reduce after backward), optional clip, optimizer step.
    Returns (loss, loss_dict)."""
    model.train()
    optimizer.zero_grad(set_to_none=True)
    with torch.autocast('cuda', dtype=torch.float16, enabled=bool(autocast)):
        loss, loss_dict = model(batch, return_loss=True)
    if scaler is not None:
        scaler.scale(loss).backward()
    else:
        loss.backward()
    if reducer is not None:
        reducer.finish()
    else:
        allreduce_grads(model, group)
    if scaler is not None:
        scaler.unscale_(optimizer)
    if grad_clip:
        torch.nn.utils.clip_grad_norm_([p for p in model.parameters() if p.requires_grad], grad_clip)
    if scaler is not None:
        scaler.step(optimizer)
        scaler.update()
    else:
        optimizer.step()
    return loss.detach(), {k: v.detach() for k, v in loss_dict.items()}
