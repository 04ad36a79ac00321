"""Multi-GPU check (run under torchrun on N GPUs of one box):
  1. whole-plot inference (BASELINE.json config 4 in miniature): overlapping synthetic tiles sharded over the ranks, one
     ragged NCCL all-gather of the inner rows, replicated merge + clustering -- must equal the single-rank result exactly;
  2. data-parallel training step (config 5 in miniature): per-rank batch, NCCL gradient all-reduce -- every rank must end
     with identical parameters, equal to a single process that averages the per-rank gradients itself.
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/dist_check.py"""
import os
import sys
from types import SimpleNamespace

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from treelearn_b200 import TreeLearn, synth  # noqa: E402
from treelearn_b200 import dist as tdist  # noqa: E402

rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
G = SimpleNamespace(tree_conf_thresh=0.5, tau_vert=0.6, tau_off=4, tau_group=0.15, tau_min=50, use_hdbscan=False)

# ---- 1. whole-plot inference -------------------------------------------------------------------------------------
plot = synth.synth_forest(edge=24.0, n_trees=30, seed=7)
tiles = []
xyz = torch.from_numpy(plot['coords'])
for cx in (-6.0, 0.0, 6.0):
    for cy in (-6.0, 0.0, 6.0):                      # 3x3 tiles of 14 m with an 8 m inner square, stride 6 m => real overlaps
        sel = ((xyz[:, 0] - cx).abs() < 7.0) & ((xyz[:, 1] - cy).abs() < 7.0)
        t = {k: (v[sel.numpy()] if hasattr(v, 'shape') and len(v) == len(xyz) else v) for k, v in plot.items()}
        t = dict(t)
        t['coords'] = (t['coords'] - [cx, cy, 0.0]).astype('float32')
        t['centre'] = (plot['centre'] + [cx, cy, 0.0]).astype('float32')
        tiles.append(synth.make_batch([t], inner_edge=8.0))
torch.manual_seed(0)
net = synth.randomize_bn_stats(TreeLearn(use_feats=False, use_coords=False, spatial_shape=[500, 500, 1000], mode='f16')).cuda().eval()
coords, labels, ncl = tdist.segment_plot(net, tiles, G)
torch.cuda.synchronize()
rows_local = coords.shape[0]
sig = torch.tensor([rows_local, int(ncl), int(labels.sum()), int((labels > 0).sum())], device='cuda', dtype=torch.int64)
if world > 1:
    sigs = [torch.zeros_like(sig) for _ in range(world)]
    dist.all_gather(sigs, sig)
    assert all(torch.equal(s, sigs[0]) for s in sigs), [s.tolist() for s in sigs]
# single-process reference: run all tiles here and compare (cheap: 9 small tiles)
import treelearn_b200.dist as tdist_mod  # noqa: E402
dist_is_init = dist.is_initialized
try:
    dist.is_initialized = lambda: False          # makes segment_plot / allgather_rows take the single-rank path
    c1, l1, n1 = tdist_mod.segment_plot(net, tiles, G)
finally:
    dist.is_initialized = dist_is_init
assert torch.equal(c1, coords) and torch.equal(l1, labels) and int(n1) == int(ncl), 'sharded plot != single-rank plot'
if rank == 0:
    print(f'plot check ok: world {world}, {len(tiles)} tiles, {coords.shape[0]} merged points, {int(ncl)} instances')

# ---- 2. data-parallel training step ------------------------------------------------------------------------------
def make_net():
    torch.manual_seed(1)
    return TreeLearn(channels=32, num_blocks=3, use_feats=False, use_coords=False, spatial_shape=[500, 500, 1000], mode='fp32').cuda()

batches = [synth.make_batch([synth.synth_forest(edge=6.0, n_trees=3, seed=20 + r, ground_density=200.0)], inner_edge=4.0)
           for r in range(world)]
net = make_net()
opt = torch.optim.SGD(net.parameters(), lr=1e-2)
loss, _ = tdist.train_step(net, opt, batches[rank])
torch.cuda.synchronize()
# reference on this rank: average of every rank's gradient, computed serially with fresh replicas
ref = make_net()
ropt = torch.optim.SGD(ref.parameters(), lr=1e-2)
acc = None
for r in range(world):
    rep = make_net().train()
    l, _ = rep(batches[r], return_loss=True)
    l.backward()
    g = [p.grad.clone() for p in rep.parameters()]
    acc = g if acc is None else [a + b for a, b in zip(acc, g)]
for p, a in zip(ref.parameters(), acc):
    p.grad = a / world
ropt.step()
err = max((p.detach() - q.detach()).abs().max().item() for p, q in zip(net.parameters(), ref.parameters()))
step = max((1e-2 * a / world).abs().max().item() for a in acc)       # largest parameter update of this step
assert err < 1e-3 * step + 1e-7, (err, step)                        # fp32 atomics order only
if rank == 0:
    print(f'DP training check ok: world {world}, max |param - reference| after one all-reduced step = {err:.2e} '
          f'(largest update {step:.2e}), loss {loss.item():.4f}')
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
