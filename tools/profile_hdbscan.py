"""Timing of the GPU HDBSCAN pieces vs sklearn on the host (run on the GPU box).  usage: python tools/profile_hdbscan.py [n ...]"""
import os
import sys
import time
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from treelearn_b200 import pipeline  # noqa: E402

warnings.simplefilter('ignore')
sizes = [int(a) for a in sys.argv[1:]] or [10000, 40000]
for n in sizes:
    rng = np.random.default_rng(n)
    k = max(n // 250, 4)
    c = rng.uniform(0, 60, (k, 2))
    pts = np.concatenate([c[i] + rng.normal(0, rng.uniform(0.04, 0.2), (200, 2)) for i in range(k)] +
                         [rng.uniform(0, 60, (n - 200 * k, 2))]).astype(np.float32)
    p = torch.from_numpy(pts).cuda()
    pipeline.hdbscan_cuda(p[:2000].contiguous(), 50)
    torch.cuda.synchronize()
    t0 = time.time()
    lab = pipeline.hdbscan_cuda(p, 50)
    torch.cuda.synchronize()
    t_gpu = time.time() - t0
    msg = f'n={len(pts)}: treelearn_b200 hdbscan {t_gpu:.3f} s ({lab.max() + 1} clusters)'
    if n <= 40000:
        from sklearn.cluster import HDBSCAN
        t0 = time.time()
        ref = HDBSCAN(min_cluster_size=50).fit_predict(pts)
        msg += f'; sklearn {time.time() - t0:.2f} s; labels identical: {np.array_equal(ref, lab)}'
    print(msg, flush=True)
