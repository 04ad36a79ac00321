"""Device timings of the post-path kernels on plot-sized synthetic inputs:
    python tools/profile_post.py            # prints one line per kernel: ms and algorithmic GB/s
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from treelearn_b200 import post  # noqa: E402
from treelearn_b200.pipeline import knn_vote_cuda  # noqa: E402


def timed(fn, reps=5):
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
    torch.manual_seed(0)
    dev = torch.device('cuda', 0)
    # voxel -> voxel join: a 0.1 m voxel cloud of a plot (4 M voxels), 90 % of them carry a prediction
    n = 4_000_000
    ret = (torch.unique(torch.randint(0, 1500, (n, 3), device=dev), dim=0).double() / 10.0).contiguous()
    lattice = ret
    sel = torch.randperm(len(lattice), device=dev)[: int(0.9 * len(lattice))]
    cur = (ret[sel] + (torch.rand((len(sel), 3), device=dev, dtype=torch.float64) - 0.5) * 0.008).contiguous()
    vals = torch.randint(0, 3000, (len(sel),), device=dev)
    ms = timed(lambda: post.hash_join_last_cuda(cur, True, vals, ret, False))
    gb = (len(sel) * 32 + len(lattice) * 32) / 1e9
    print(f'hash join  build {len(sel)} rows, probe {len(lattice)} rows (fp64): {ms:8.3f} ms  {gb / ms * 1e3:7.1f} GB/s algorithmic')
    # co-occurrence counts of a 20 M-point plot with ~2000 trees
    m = 20_000_000
    pred = torch.sort(torch.randint(-1, 2000, (m,), device=dev)).values
    gt = (pred + torch.randint(-1, 2, (m,), device=dev)).clamp_(-1, 1999)
    ms = timed(lambda: post.cooccurrence_counts_cuda(pred, gt, 2000, 2000))
    print(f'co-occurrence counts {m} points, 2001 x 2001 table: {ms:8.3f} ms  {m * 16 / ms / 1e6:7.1f} GB/s algorithmic')
    # kNN(5) propagation: 2 M voxel predictions -> 8 M original points
    src = torch.rand((2_000_000, 3), device=dev) * torch.tensor([60.0, 60.0, 30.0], device=dev)
    lab = (src[:, 0] / 3).long() + 20 * (src[:, 1] / 3).long()
    tgt = (src[torch.randint(0, len(src), (8_000_000,), device=dev)] + torch.randn((8_000_000, 3), device=dev) * 0.03).contiguous()
    ms = timed(lambda: knn_vote_cuda(src, lab, tgt, 5), reps=2)
    print(f'propagate_preds kNN(5) {len(src)} sources -> {len(tgt)} targets: {ms:8.3f} ms  {len(tgt) / ms / 1e3:7.2f} M targets/s')


if __name__ == '__main__':
    main()
