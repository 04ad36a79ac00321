"""A/B harness for the experimental TS-form conv kernel (make_ts_conv.py): routes every fp16 `tl_conv_fwd` call of the
product's Python path to tools/experiments/_build/libtl_conv_ts.so and compares results and per-layer times with the
product library.  Needs a B200.  NOT RUN YET (written at the end of round 1).

    python tools/experiments/ts_conv_check.py [workload=cfg2_2M]

Order of business in round 2: (1) ts_mma_probe.py, (2) this script with TL_TS_MAX_N=32 on the `small` workload (parity),
(3) cfg2_2M timing, (4) only then move the code into treelearn_b200/csrc/tl_conv_tc.cu behind the existing tests.
Wrap the call in `timeout 120`: a barrier mistake in the new producer path hangs the kernel.
"""
import collections
import ctypes as C
import os
import subprocess
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from treelearn_b200 import TreeLearn, _lib, sparse, synth  # noqa: E402

SO = os.path.join(HERE, '_build', 'libtl_conv_ts.so')
USE_TS = [False]


def install_router():
    if not os.path.exists(SO):
        subprocess.run([sys.executable, os.path.join(HERE, 'make_ts_conv.py')], check=True)
    exp = C.CDLL(SO)
    exp.tl_conv_fwd_ts.restype = C.c_int
    exp.tl_conv_fwd_ts.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    exp.tl_last_error_ts.restype = C.c_char_p
    lib = _lib.load()
    product = lib.tl_conv_fwd

    def routed(desc_ref, mode, stream):
        if USE_TS[0] and mode == _lib.MODE_F16:
            rc = exp.tl_conv_fwd_ts(desc_ref, mode, stream)
            if rc != 0:
                print('experimental library:', exp.tl_last_error_ts().decode())
            return rc
        return product(desc_ref, mode, stream)

    lib.tl_conv_fwd = routed


def forward(net, dev, profile=False):
    with torch.no_grad():
        for _ in range(2):
            out = net(dev, return_loss=False)
        torch.cuda.synchronize()
        layers = None
        if profile:
            sparse.PROFILE = []
            out = net(dev, return_loss=False)
            torch.cuda.synchronize()
            layers = collections.OrderedDict()
            for e0, e1, byts, flops, c_out in sparse.PROFILE:
                layers.setdefault(c_out, [0, 0.0])
                layers[c_out][0] += 1
                layers[c_out][1] += e0.elapsed_time(e1)
            sparse.PROFILE = None
    return out, layers


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else 'cfg2_2M'
    install_router()
    for name in ('small', workload):
        batch = synth.make_batch([synth.workload(name)])
        dev = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()
               if k in ('coords', 'input_feats', 'batch_ids', 'batch_size')}
        net = synth.randomize_bn_stats(TreeLearn(use_feats=False, use_coords=False, spatial_shape=[1000, 1000, 1000],
                                                 mode='f16')).cuda().eval()
        USE_TS[0] = False
        ref, ref_layers = forward(net, dev, profile=True)
        USE_TS[0] = True
        got, got_layers = forward(net, dev, profile=True)
        print(f'== {name}: {dev["coords"].shape[0]} points')
        for k in ref:
            err = (ref[k].float() - got[k].float()).abs().max().item()
            print(f'   {k:28s} max |TS - product| = {err:.3e}')      # same arithmetic (fp16 operands, fp32 accumulate): expect ~1e-6
        print('   conv ms by C_out (launches):  product -> TS')
        for c_out in ref_layers:
            print(f'   C_out {c_out:4d} ({ref_layers[c_out][0]:2d}): {ref_layers[c_out][1]:8.3f} -> {got_layers[c_out][1]:8.3f}')
        print(f'   total {sum(v[1] for v in ref_layers.values()):8.3f} -> {sum(v[1] for v in got_layers.values()):8.3f} ms')


if __name__ == '__main__':
    main()
