"""Per-layer device time of the conv schedule (CUDA events around every launch) -- run on the GPU box.
usage: python tools/profile_layers.py [workload] [mode]"""
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from treelearn_b200 import TreeLearn, synth, sparse  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else 'cfg1_200k'
mode = sys.argv[2] if len(sys.argv) > 2 else 'tf32'
shape = [1000, 1000, 1000]
batch = synth.make_batch([synth.workload(workload)])
dev = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items() if k in ('coords', 'input_feats', 'batch_ids', 'batch_size')}
net = synth.randomize_bn_stats(TreeLearn(use_feats=False, use_coords=False, spatial_shape=shape, mode=mode)).cuda().eval()
ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
with torch.no_grad():
    for _ in range(3):
        net(dev, return_loss=False)
    torch.cuda.synchronize()
    sparse.PROFILE = []
    t = [ev() for _ in range(5)]
    t[0].record()
    vf, vc, keys, v2p = sparse.voxelize(dev['coords'], dev['input_feats'], dev['batch_ids'], 1, 0.1, False, False, 3)
    t[1].record()
    levels = sparse.build_levels(keys, vc, shape, 7)
    t[2].record()
    out = net._run_backbone(vf, levels)
    t[3].record()
    net.forward_head(out, v2p)
    t[4].record()
    torch.cuda.synchronize()
names = ['voxelize', 'levels+rulebooks', 'backbone convs', 'heads']
for i, n in enumerate(names):
    print(f'{n:20s} {t[i].elapsed_time(t[i + 1]):8.3f} ms')
print('voxels per level:', [lv.n for lv in levels])
agg = collections.OrderedDict()
for e0, e1, byts, flops, c_out in sparse.PROFILE:
    key = (c_out,)
    a = agg.setdefault(key, [0, 0.0, 0, 0])
    a[0] += 1
    a[1] += e0.elapsed_time(e1)
    a[2] += byts
    a[3] += flops
print(f'{"c_out":>6} {"launches":>8} {"ms":>9} {"GB/s(alg)":>10} {"denseTF/s":>10}')
for (c,), (n, ms, b, f) in agg.items():
    print(f'{c:6d} {n:8d} {ms:9.3f} {b / ms / 1e6:10.1f} {f / ms / 1e9:10.2f}')
print('total conv ms', sum(a[1] for a in agg.values()))
for i, (e0, e1, byts, flops, c_out) in enumerate(sparse.PROFILE):
    print(i, c_out, round(e0.elapsed_time(e1), 4), byts)
