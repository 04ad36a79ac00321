"""Training step timing (BASELINE.json config 3: 4-tile batch, fwd + loss + bwd + AdamW step) -- run on the GPU box.
usage: python tools/profile_train.py [n_tiles] [mode] [steps]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from treelearn_b200 import TreeLearn, synth, sparse  # noqa: E402

n_tiles = int(sys.argv[1]) if len(sys.argv) > 1 else 4
mode = sys.argv[2] if len(sys.argv) > 2 else 'tf32'
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
tiles = [synth.synth_forest(edge=20.0, n_trees=20, seed=s) for s in range(n_tiles)]
batch = synth.make_batch(tiles)
torch.manual_seed(0)
net = TreeLearn(use_feats=False, use_coords=False, spatial_shape=[500, 500, 1000], mode=mode).cuda().train()
opt = torch.optim.AdamW(net.parameters(), lr=2e-3, weight_decay=1e-3)
dev = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
with torch.no_grad():
    _, vc, _, _ = sparse.voxelize(dev['coords'], dev['input_feats'], dev['batch_ids'], n_tiles, 0.1, False, False, 3)
n_vox = vc.shape[0]
ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
for it in range(steps + 1):
    e = [ev() for _ in range(4)]
    e[0].record()
    loss, ld = net(dev, return_loss=True)
    e[1].record()
    opt.zero_grad(set_to_none=True)
    loss.backward()
    e[2].record()
    opt.step()
    e[3].record()
    torch.cuda.synchronize()
    f, b, o = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3])
    print(f'step {it}: loss {loss.item():.4f}  fwd {f:8.2f} ms  bwd {b:8.2f} ms  opt {o:6.2f} ms  total {f + b + o:8.2f} ms  '
          f'{n_vox / (f + b + o) / 1e3:8.2f} Mvoxels/s  ({n_vox} voxels, {n_tiles} tiles, mode {mode}, '
          f'peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB)')
