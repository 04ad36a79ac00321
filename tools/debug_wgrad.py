"""Debug harness of csrc/tl_wgrad_tc.cu: dW against a torch einsum on a small level, for descriptor variants given by the
TL_WG_* environment variables (one process per variant: the library reads them per call).  Run on the GPU box."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from treelearn_b200 import _lib, sparse, synth  # noqa: E402

ci, co = int(sys.argv[1]) if len(sys.argv) > 1 else 32, int(sys.argv[2]) if len(sys.argv) > 2 else 32
batch = synth.make_batch([synth.synth_forest(edge=4.0, n_trees=2, seed=10, ground_density=150.0)])
vf, vc, keys, v2p = sparse.voxelize(batch['coords'].cuda(), batch['input_feats'].cuda(), batch['batch_ids'].cuda(), 1, 0.1, False, True, 3)
lv = sparse.build_levels(keys, vc, [500, 500, 1000], 1)[0]
g = torch.Generator().manual_seed(1)
x = torch.randn((lv.n, ci), generator=g).cuda()
gy = torch.randn((lv.n, co), generator=g).cuda()
nbr = lv.nbr[:, :lv.n].long()
ref = torch.zeros((27, ci, co), device='cuda')
for k in range(27):
    m = nbr[k] >= 0
    ref[k] = x[nbr[k][m]].T @ gy[m]
lib = _lib.load()
dw = torch.full((27, ci, co), float('nan'), device='cuda')
wsb = lib.tl_conv_wgrad_tc_workspace_bytes(lv.n, ci, 27, co)
ws = torch.empty(max(int(wsb), 256), dtype=torch.uint8, device='cuda')
_lib.check(lib.tl_conv_wgrad_tc(_lib.ptr(x), ci, ci, 27, _lib.ptr(lv.nbr), lv.nbr.stride(0), _lib.ptr(lv.nbr_mask), _lib.ptr(gy), lv.n, co,
                                _lib.ptr(dw), _lib.ptr(ws), wsb, _lib.stream_ptr()))
torch.cuda.synchronize()
err = (dw - ref).abs()
print(f'variant {dict((k, v) for k, v in os.environ.items() if k.startswith("TL_WG_"))}: n={lv.n} ci={ci} co={co} max|ref|={ref.abs().max():.3f} '
      f'max|dw|={dw.abs().max():.3f} zeros={float((dw == 0).float().mean()):.3f} nan={float(dw.isnan().float().mean()):.3f} '
      f'max err={err.max():.4f} mean err={err.mean():.4f}')
for k in (0, 13, 26):
    print('  k', k, 'err', float(err[k].max()), 'dw[0,:4]', dw[k, 0, :4].tolist(), 'ref[0,:4]', ref[k, 0, :4].tolist())
# which permutation of rows / columns matches? correlation of dw[13] against ref[13] and its transposes
a = dw[13].flatten()
for name, b in (('ref', ref[13]), ('ref.T', ref[13].T if ci == co else None)):
    if b is not None:
        b = b.flatten()
        print('  corr(dw[13],', name, ') =', float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-30)))
