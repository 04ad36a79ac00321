"""Timeline of the k_conv_tc ring on SM 0 (TL_TC_DEBUG bit 32) for one level-0 32->32 conv of the cfg2 tile:
per slot fill: gather wait / issue, data latency, MMA issue, slot-free latency.  Run on the GPU box:
    make -C treelearn_b200/csrc clean all TRACE=1 && TL_TC_DEBUG=32 python tools/trace_conv.py [32|64]
(the hooks are compiled out of the default build)"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from treelearn_b200 import _lib, sparse, synth  # noqa: E402

assert int(os.environ.get('TL_TC_DEBUG', '0')) & 32, 'run with TL_TC_DEBUG=32'
c_ch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
level = {32: 0, 64: 1, 96: 2}[c_ch]
batch = synth.make_batch([synth.workload('cfg2_2M')])
dev = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
vf, vc, keys, v2p = sparse.voxelize(dev['coords'], dev['input_feats'], dev['batch_ids'], 1, 0.1, False, False, 3)
lv = sparse.build_levels(keys, vc, [1000, 1000, 1000], level + 1)[level]
g = torch.Generator(device='cuda').manual_seed(0)
x = torch.randn((lv.n, c_ch), device='cuda', generator=g).half()
w = sparse.pack_weight_tc(torch.randn((27, c_ch, c_ch), device='cuda', generator=g) / 30, True)
s, t = torch.ones(c_ch, device='cuda'), torch.zeros(c_ch, device='cuda')
for _ in range(3):
    out = sparse.conv([sparse.Seg(x, w, lv.nbr, lv.nbr_mask)], lv.n, c_ch, _lib.MODE_F16, act1=(s, t))
torch.cuda.synchronize()
lib = C.CDLL(_lib.LIB_PATH)
ROLES, LEN = 8, 2048
buf = np.zeros(ROLES * LEN, dtype=np.uint64)
assert lib.tl_debug_copy_trace(C.c_void_p(buf.ctypes.data), C.c_size_t(buf.nbytes)) == 0
buf = buf.reshape(ROLES, LEN)
tag, clk = (buf >> np.uint64(48)).astype(np.int64), (buf & np.uint64((1 << 48) - 1)).astype(np.int64)
ev = {}
for role in range(ROLES):
    for p in range(LEN):
        if buf[role, p] == 0:
            break
        ev[(role, int(tag[role, p]) >> 2, int(tag[role, p]) & 3)] = int(clk[role, p])
groups = sorted({r for (r, _, _) in ev if r < 4})
fills = sorted({f for (r, f, _) in ev if r == 4})
print(f'traced fills: {len(fills)}; producer groups: {groups}')
rows = []
for f in fills[20:400]:
    grp = [r for r in groups if (r, f, 2) in ev]
    if not grp or (4, f, 2) not in ev:
        continue
    r = grp[0]
    rows.append((ev[(r, f, 1)] - ev[(r, f, 0)],          # gather group waits for its slot
                 ev[(r, f, 2)] - ev[(r, f, 1)],          # issue of the copies (lane 0 of warp 0)
                 ev[(4, f, 1)] - ev[(r, f, 2)],          # copies issued -> MMA warp sees the slot full (data latency + hop)
                 ev[(4, f, 1)] - ev[(4, f, 0)],          # MMA warp's wait for this slot
                 ev[(4, f, 2)] - ev[(4, f, 1)],          # MMA issue + commit
                 ev[(4, f, 3)] - ev[(4, f, 1)] if (4, f, 3) in ev else 0,   # ... of which the MMAs
                 (ev[(4, f + 1, 0)] - ev[(4, f, 2)]) if (4, f + 1, 0) in ev else 0))   # loop overhead until the next wait
a = np.array(rows, dtype=np.float64)
names = ['gather: wait for free slot', 'gather: issue copies', 'issued -> full seen by MMA', 'MMA: wait for full', 'MMA: issue + commit',
         'MMA: fence + elect + MMAs only', 'MMA: commit done -> next wait begins']
for i, n in enumerate(names):
    print(f'{n:32s} median {np.median(a[:, i]):8.0f}  mean {a[:, i].mean():8.0f}  p90 {np.percentile(a[:, i], 90):8.0f} cycles')
per_fill = (ev[(4, fills[400], 2)] - ev[(4, fills[20], 2)]) / 380 if len(fills) > 400 else float('nan')
print(f'MMA warp: {per_fill:.0f} cycles per fill in steady state')
# slot-free latency: MMA commit of fill f  ->  the gather group that reuses the slot sees it free
free = []
for f in fills[20:400]:
    for r in groups:
        for f2 in range(f + 1, f + 8):
            if (r, f2, 1) in ev and (r, f2, 0) in ev and (4, f, 2) in ev and ev[(r, f2, 1)] > ev[(4, f, 2)] and ev[(r, f2, 0)] < ev[(4, f, 2)]:
                free.append(ev[(r, f2, 1)] - ev[(4, f, 2)])
if free:
    print(f'commit -> slot seen free by a waiting gather group: median {np.median(free):.0f} cycles ({len(free)} samples)')
