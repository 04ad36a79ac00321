"""Round-2 probe (see ts_mma_probe.cu): time M128 x N x K16 fp16 tcgen05.mma with A in shared memory vs A in TMEM and check
both against A @ B^T.  Needs a B200:

    python tools/experiments/ts_mma_probe.py            # builds ts_mma_probe.so next to this file, then runs

Expected if the A read is what bounds the small-N MMAs of k_conv_tc: SS ~64 cycles per MMA for every N <= 128,
TS ~N/2 cycles (16 at N = 32).  If the TS result is wrong, the assumed TMEM layout of A (row m in lane m, element k in
column k/2, low half first) is the first thing to check.
"""
import ctypes as C
import os
import subprocess

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, 'ts_mma_probe.so')


def main():
    if not os.path.exists(SO):
        subprocess.run(['nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-shared', '-Xcompiler', '-fPIC', '-o', SO,
                        os.path.join(HERE, 'ts_mma_probe.cu')], check=True)
    lib = C.CDLL(SO)
    lib.ts_mma_probe.restype = C.c_int
    lib.ts_mma_probe.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    torch.manual_seed(0)
    chunks, reps = 8, 64
    for n in (32, 64, 128, 256):
        a = (torch.randn(128, 32 * chunks, device='cuda') * 0.25).half()
        b = (torch.randn(n, 32 * chunks, device='cuda') * 0.25).half()
        d_ss, d_ts = torch.zeros(128, n, device='cuda'), torch.zeros(128, n, device='cuda')
        cyc = torch.zeros(2, dtype=torch.int64, device='cuda')
        rc = lib.ts_mma_probe(a.data_ptr(), b.data_ptr(), n, chunks, reps, d_ss.data_ptr(), d_ts.data_ptr(), cyc.data_ptr(), None)
        torch.cuda.synchronize()
        assert rc == 0, rc
        ref = a.float() @ b.float().t()
        mmas = reps * chunks * 2
        print(f'N={n:3d}: SS {cyc[0].item() / mmas:6.1f} cycles/MMA (max err {float((d_ss - ref).abs().max()):.2e})   '
              f'TS {cyc[1].item() / mmas:6.1f} cycles/MMA (max err {float((d_ts - ref).abs().max()):.2e})   floor N/2 = {n // 2}')


if __name__ == '__main__':
    main()
