"""Device timings of the pre-path kernels on a plot-sized synthetic cloud:
    python tools/profile_prepare.py [n_raw_points]
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from treelearn_b200 import prepare, synth  # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    n_raw = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
    dev = torch.device('cuda', 0)
    f = synth.synth_forest(edge=60.0, height=20.0, n_trees=150, seed=0, ground_density=300.0)
    base = torch.from_numpy(f['coords'].astype(np.float64)).to(dev)
    torch.manual_seed(0)
    pick = torch.randint(0, len(base), (n_raw,), device=dev)
    raw = (base[pick] + torch.randn((n_raw, 3), device=dev, dtype=torch.float64) * 0.04 + 500.0).contiguous()
    bound = float(raw.abs().max()) + 100
    ms = timed(lambda: prepare.voxel_downsample_trace_cuda(raw, 0.1, -bound - 0.05, round2_first=True))
    down = prepare.voxel_downsample_trace_cuda(raw, 0.1, -bound - 0.05, round2_first=True)[0]
    print(f'voxel down-sample + trace: {n_raw} points -> {len(down)} voxels: {ms:8.2f} ms  '
          f'({n_raw * (24 + 8) / ms / 1e6:6.1f} GB/s algorithmic, {n_raw / ms / 1e3:6.1f} M points/s)')
    pts = (torch.round(down.float() * 100) / 100).double().contiguous()
    ms = timed(lambda: prepare.verticality_cuda(pts, 0.6), reps=2)
    print(f'verticality (radius 0.6 m): {len(pts)} voxels: {ms:8.2f} ms  ({len(pts) / ms / 1e3:6.2f} M points/s)')
    feats = prepare.compute_features(pts.cpu().numpy(), 0.6)
    plot = pts.float().cpu().numpy()
    t0 = time.time()
    tiles = prepare.cut_tiles(plot, np.zeros(len(plot), np.float32), feats, 8, 13.5, 0.5)
    torch.cuda.synchronize()
    print(f'tile cutting (8 m inner, 13.5 m context, stride 0.5): {len(tiles)} tiles, {sum(len(t["points"]) for t in tiles)} rows: '
          f'{(time.time() - t0) * 1e3:8.1f} ms wall (incl. D2H of every tile)')


if __name__ == '__main__':
    main()
