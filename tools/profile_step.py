"""Device time of every phase of one bench step (CUDA events) -- run on the GPU box.
usage: python tools/profile_step.py [workload] [mode]"""
import os
import sys
from types import SimpleNamespace

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from treelearn_b200 import TreeLearn, synth, sparse, pipeline  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else 'cfg2_2M'
mode = sys.argv[2] if len(sys.argv) > 2 else 'f16'
G = SimpleNamespace(tree_conf_thresh=0.5, tau_vert=0.6, tau_off=4, tau_group=0.15, tau_min=50, use_hdbscan=False)
shape = [1000, 1000, 1000]
batch = synth.make_batch([synth.workload(workload)])
dev = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items() if k in ('coords', 'input_feats', 'batch_ids', 'batch_size')}
torch.manual_seed(0)   # same random-init network as bench.py
net = synth.randomize_bn_stats(TreeLearn(use_feats=False, use_coords=False, spatial_shape=shape, mode=mode)).cuda().eval()


class T:
    def __init__(self):
        self.marks = []

    def mark(self, name):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        self.marks.append((name, e))

    def report(self):
        torch.cuda.synchronize()
        for (n0, e0), (n1, e1) in zip(self.marks[:-1], self.marks[1:]):
            print(f'{n1:34s} {e0.elapsed_time(e1):9.3f} ms')
        print(f'{"TOTAL":34s} {self.marks[0][1].elapsed_time(self.marks[-1][1]):9.3f} ms')


with torch.no_grad():
    for it in range(3):
        t = T()
        t.mark('start')
        vf, vc, keys, v2p = sparse.voxelize(dev['coords'], dev['input_feats'], dev['batch_ids'], 1, 0.1, False, False, 3)
        t.mark('voxelize')
        levels = sparse.build_levels(keys, vc, shape, 7, subm=False)
        t.mark('level maps')
        for lv in levels:
            sparse.build_subm_rulebook(lv)
        t.mark('subm rulebooks')
        out = net._run_backbone(vf, levels)
        t.mark('backbone convs')
        o = net.forward_head(out, v2p)
        t.mark('heads')
        coords, offs, logits, vert = dev['coords'], o['offset_predictions'], o['semantic_prediction_logits'], dev['input_feats'][:, -1]
        shifted = coords + offs
        tree_mask = logits.float().softmax(dim=-1)[:, 0] >= 0.5
        mask = tree_mask & (vert > 0.6) & (offs[:, 2].abs() < 4)
        ind = mask.nonzero().squeeze(1)
        pred = torch.full((coords.shape[0],), 0, dtype=torch.int64, device=coords.device)
        pred[tree_mask] = -1
        pts = shifted[ind][:, :2].contiguous()
        t.mark('masks (torch)')
        lab, ncl = pipeline.group_dbscan_cuda(pts, 0.15, 50, -1, 1)
        t.mark(f'cluster cc ({pts.shape[0]} pts -> {ncl})')
        pred[ind] = lab
        tree_idx = (pred != 0).nonzero().squeeze(1)
        tp = pred[tree_idx]
        q = (tp == -1).nonzero().squeeze(1)
        r = (tp != -1).nonzero().squeeze(1)
        sh = shifted[tree_idx]
        a, b_, c = sh[r].contiguous(), tp[r].contiguous(), sh[q].contiguous()
        t.mark('knn prep (torch)')
        if q.numel() and r.numel() >= 5:
            tp[q] = pipeline.knn_vote_cuda(a, b_, c, 5)
        t.mark(f'knn vote ({r.numel()} ref, {q.numel()} query)')
        if it == 2:
            t.report()
