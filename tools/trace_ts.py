"""Timeline of the f16 conv kernel on SM 0 (TRACE build, TL_GRP_DEBUG / TL_TS_DEBUG bit 32) for one submanifold conv of the cfg2 tile.
    make -C treelearn_b200/csrc TRACE=1
    TL_LIB=treelearn_b200/libtreelearn_b200_trace.so TL_GRP_DEBUG=32 python tools/trace_ts.py [32|64|96|128] [f16|f16x2]
    (TL_TS=2 TL_TS_DEBUG=32 ... for the tensor-memory-A kernel)
Events (tag = tile ordinal << 3 | event) of the four warps of group 0 on CTA 0 (roles 0..3):
  0 tile begins, 1 main loop done, 2 accumulator complete, 3 epilogue done"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from treelearn_b200 import _lib, sparse, synth  # noqa: E402

KIND = 'grp' if os.environ.get('TL_TS', '1') == '1' else 'ts'     # which kernel mode f16 dispatches to (csrc/tl_conv_simt.cu)
assert int(os.environ.get(f'TL_{KIND.upper()}_DEBUG', '0')) & 32, f'run with TL_{KIND.upper()}_DEBUG=32 and the TRACE build (TL_LIB=...)'
c_ch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
mode = sys.argv[2] if len(sys.argv) > 2 else 'f16'
nsplit = 2 if mode == 'f16x2' else 1
level = {32: 0, 64: 1, 96: 2, 128: 3}[c_ch]
batch = synth.make_batch([synth.workload('cfg2_2M')])
dev = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
vf, vc, keys, v2p = sparse.voxelize(dev['coords'], dev['input_feats'], dev['batch_ids'], 1, 0.1, False, False, 3)
lv = sparse.build_levels(keys, vc, [1000, 1000, 1000], level + 1)[level]
g = torch.Generator(device='cuda').manual_seed(0)
x = torch.randn((lv.n, c_ch), device='cuda', generator=g)
x = sparse.to_split(x) if nsplit == 2 else x.half()
w = sparse.pack_weight(torch.randn((27, c_ch, c_ch), device='cuda', generator=g) / 30, nsplit)
s, t = torch.ones(c_ch, device='cuda'), torch.zeros(c_ch, device='cuda')
m = _lib.MODE_F16X2 if nsplit == 2 else _lib.MODE_F16
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    e0.record()
    out = sparse.conv([sparse.Seg(x, w, lv.nbr, lv.nbr_mask)], lv.n, c_ch, m, act1=(s, t))
    e1.record()
torch.cuda.synchronize()
tiles = (lv.n + 127) // 128
print(f'level {level}: {lv.n} voxels, {tiles} tiles ({tiles / 148:.1f} per SM), C = {c_ch}, mode {mode}: {e0.elapsed_time(e1) * 1e3:.1f} us')
lib = C.CDLL(_lib.LIB_PATH)
ROLES, LEN = 8, 4096
buf = np.zeros(ROLES * LEN, dtype=np.uint64)
assert getattr(lib, f'tl_debug_copy_trace_{KIND}')(C.c_void_p(buf.ctypes.data), C.c_size_t(buf.nbytes)) == 0
buf = buf.reshape(ROLES, LEN)
tag, clk = (buf >> np.uint64(48)).astype(np.int64), (buf & np.uint64((1 << 48) - 1)).astype(np.int64)
ev = {}
for role in range(ROLES):
    for p in range(LEN):
        if buf[role, p] == 0:
            break
        ev[(role, int(tag[role, p]) >> 3, int(tag[role, p]) & 7)] = int(clk[role, p])


def stat(name, vals):
    a = np.array(vals, dtype=np.float64)
    if len(a):
        print(f'{name:58s} median {np.median(a):7.0f}  mean {a.mean():7.0f}  p90 {np.percentile(a, 90):7.0f}  n {len(a)}')


# roles 0..3 = the four warps (TMEM lane quarters) of group 0 on CTA 0; per round (= tile): 0 begin, 1 main loop done,
# 2 accumulator complete seen, 3 epilogue done
rounds = sorted({r for (role, r, e) in ev if role == 0 and e == 3})
print(f'traced: {len(rounds)} tiles of group 0 on SM 0')
R_ = rounds[1:-1]
for w in range(4):
    stat(f'warp {w}: main loop (gather -> TMEM -> MMA issue)', [ev[(w, r, 1)] - ev[(w, r, 0)] for r in R_ if (w, r, 1) in ev])
    stat(f'warp {w}: next-tile index staging + wait for accumulator', [ev[(w, r, 2)] - ev[(w, r, 1)] for r in R_ if (w, r, 2) in ev])
    stat(f'warp {w}: epilogue', [ev[(w, r, 3)] - ev[(w, r, 2)] for r in R_ if (w, r, 3) in ev])
    stat(f'warp {w}: epilogue done -> next tile begins', [ev[(w, r + 1, 0)] - ev[(w, r, 3)] for r in R_ if (w, r + 1, 0) in ev])
if len(R_) > 2:
    print(f'group 0: {(ev[(0, R_[-1], 3)] - ev[(0, R_[0], 3)]) / (len(R_) - 1):.0f} cycles per tile in steady state')

# ---- per-chunk events (group kernel): role 4 = the group's MMA warp (0 begin wait full, 1 full seen, 2 MMAs + commit issued),
#      role 5 = gather warp 1 (0 begin wait for a free stage, 1 stage free, 2 copies issued + arrive)
def seq(role):
    out = []
    for p in range(LEN):
        if buf[role, p] == 0:
            break
        out.append((int(tag[role, p]) & 7, int(tag[role, p]) >> 3, int(clk[role, p])))
    return out


def deltas(events, a, b):
    return [events[k + 1][2] - events[k][2] for k in range(len(events) - 1) if events[k][0] == a and events[k + 1][0] == b]


m, gth = seq(4), seq(5)
if m:
    m = m[60:]
    stat('MMA warp per chunk: wait for full', deltas(m, 0, 1))
    stat('MMA warp per chunk: fence + MMAs + commit', deltas(m, 1, 2))
    stat('MMA warp per chunk: commit -> next wait (same tile)', [d for d, e in zip(deltas(m, 2, 0), [x for x in m if x[0] == 2]) if True])
    per = [m[k + 3][2] - m[k][2] for k in range(0, len(m) - 3) if m[k][0] == 0 and m[k + 3][0] == 0 and m[k + 3][1] == m[k][1] + 1]
    stat('MMA warp: chunk to chunk (same tile)', per)
if gth:
    gth = gth[30:]
    stat('gather warp 1 per own chunk: wait for a free stage', deltas(gth, 0, 1))
    stat('gather warp 1 per own chunk: address + 16 copies + arrive', deltas(gth, 1, 2))
    stat('gather warp 1: arrive -> next own chunk begins (incl. index loads)', deltas(gth, 2, 0))
    # stage latency: copies issued by the gather warp -> the MMA warp sees the chunk full
    mm = {}
    tile_no = 0
    for ev, i, c in seq(4):
        if ev == 0 and i == 0:
            tile_no += 1
        mm[(tile_no, i, ev)] = c
    tile_no, lat, lat2 = 0, [], []
    prev_i = 1 << 30
    for ev, i, c in seq(5):
        if ev == 0 and i < prev_i:
            tile_no += 1
        if ev == 0:
            prev_i = i
        if ev == 2 and (tile_no, i, 1) in mm:
            lat.append(mm[(tile_no, i, 1)] - c)
        if ev == 1 and (tile_no, i, 2) in mm:
            pass
    stat('copies issued -> MMA warp sees the chunk full', lat[10:])
