"""Diagnostics for the training path (run on the GPU box): per-parameter gradient error vs the oracle, run-to-run
difference of two identical steps, and BN+ReLU errors vs a float64 torch reference."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import model_ref  # noqa: E402
from treelearn_b200 import TreeLearn  # noqa: E402
from treelearn_b200 import autograd as ag  # noqa: E402

g = np.load(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', 'model_small.npz'))
sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith('sd:')}
batch = {k[6:]: torch.from_numpy(g[k]) for k in g.files if k.startswith('batch:')}
batch['batch_size'] = int(batch['batch_size'])
sd_ref = {k: (v.clone().requires_grad_() if v.is_floating_point() and 'running' not in k else v.clone()) for k, v in sd.items()}
out = model_ref.forward_ref(sd_ref, batch, use_coords=False, use_feats=True, spatial_shape=[500, 500, 1000], training=True, new_stats={})
loss_ref, _ = model_ref.loss_ref(out, batch)
loss_ref.backward()
runs = []
for rep in range(3):
    net = TreeLearn(channels=8, num_blocks=3, use_feats=True, use_coords=False, spatial_shape=[500, 500, 1000])
    net.load_state_dict(sd, strict=True)
    net = net.cuda().train()
    loss, _ = net(batch, return_loss=True)
    loss.backward()
    runs.append({n: p.grad.detach().cpu().clone() for n, p in net.named_parameters()})
    print('rep', rep, 'loss', loss.item(), 'ref', loss_ref.item())
for n in runs[0]:
    ref = sd_ref[n].grad
    e = [(r[n] - ref).abs().max().item() for r in runs]
    d01 = (runs[0][n] - runs[1][n]).abs().max().item()
    print(f'{n:60s} scale {ref.abs().max().item():.3e} err {e[0]:.3e} {e[1]:.3e} {e[2]:.3e}  run0-run1 {d01:.3e}')

for n, c in [(5000, 32), (100000, 224)]:
    gen = torch.Generator().manual_seed(n + c)
    x = torch.randn((n, c), generator=gen) * 2 + 0.5
    gy = torch.randn((n, c), generator=gen)
    bn64 = torch.nn.BatchNorm1d(c, eps=1e-4, momentum=0.1).double()
    with torch.no_grad():
        bn64.weight.copy_(torch.rand(c, generator=gen) + 0.5)
        bn64.bias.copy_(torch.randn(c, generator=gen) * 0.2)
    bn = torch.nn.BatchNorm1d(c, eps=1e-4, momentum=0.1)
    bn.load_state_dict({k: v.float() if v.is_floating_point() else v for k, v in bn64.state_dict().items()})
    bn = bn.cuda().train()
    x64 = x.double().requires_grad_()
    o64 = F.relu(bn64(x64))
    o64.backward(gy.double())
    xc = x.cuda().requires_grad_()
    o = ag.bn_relu(xc, bn)
    o.backward(gy.cuda())
    print(f'bn n={n} c={c}: out err {(o.detach().cpu().double() - o64.detach()).abs().max().item():.3e} '
          f'dx err {(xc.grad.cpu().double() - x64.grad).abs().max().item():.3e} '
          f'dgamma err {(bn.weight.grad.cpu().double() - bn64.weight.grad).abs().max().item():.3e} (scale {bn64.weight.grad.abs().max().item():.2e}) '
          f'dbeta err {(bn.bias.grad.cpu().double() - bn64.bias.grad).abs().max().item():.3e}')
    # fp32 torch reference on CPU for comparison of ITS error
    bn32 = torch.nn.BatchNorm1d(c, eps=1e-4, momentum=0.1)
    bn32.load_state_dict({k: v.float() if v.is_floating_point() else v for k, v in bn64.state_dict().items()})
    x32 = x.clone().requires_grad_()
    F.relu(bn32(x32)).backward(gy)
    print(f'   torch fp32 CPU dx err vs fp64 {(x32.grad.double() - x64.grad).abs().max().item():.3e}')
