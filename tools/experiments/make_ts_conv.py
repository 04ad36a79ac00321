"""Derive the experimental "A operand in TMEM" variant of k_conv_tc (TS_CONV_PLAN.md) from the product source.

    python tools/experiments/make_ts_conv.py        # writes tools/experiments/_build/tl_conv_tc_ts.cu and builds
                                                    # tools/experiments/_build/libtl_conv_ts.so (exports tl_conv_fwd_ts)

The product file stays untouched; every edit below is an exact-match replacement that fails loudly when the product
source has moved on.  NOT RUN ON HARDWARE YET (end of round 1: compile-checked only) — see ts_conv_check.py for the A/B
harness to run first in round 2.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(ROOT, 'treelearn_b200', 'csrc', 'tl_conv_tc.cu')
OUT_DIR = os.path.join(HERE, '_build')


def sub(text, old, new, count=1):
    assert text.count(old) == count, f'expected {count} occurrence(s), found {text.count(old)}:\n{old[:200]}'
    return text.replace(old, new)


TS_HELPERS = r'''
// ---- TS form (A operand in TMEM) --------------------------------------------------------------------------------
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ uint4 ld_shared_u4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
// 16 consecutive TMEM columns of this thread's lane (one 32-channel fp16 chunk row = 64 B)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint4& a, const uint4& b, const uint4& c, const uint4& d) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w), "r"(c.x), "r"(c.y), "r"(c.z), "r"(c.w),
        "r"(d.x), "r"(d.y), "r"(d.z), "r"(d.w)
        : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
'''

TS_PRODUCER = r'''    } else if (P.ts && warp < 4 * P.groups) {
        // ===================== TS producers: gather -> private staging -> registers -> TMEM A ring ==================
        // Warp r = warp % 4 of a group owns tile rows [32 r, 32 r + 32) (its TMEM lane quarter) of EVERY chunk of the
        // group's fills.  It gathers them with 16 B cp.async into two private staging buffers of QB chunks (nobody else
        // reads them: no mbarrier on the A side of shared memory, only cp.async groups + __syncwarp), reads them back
        // one row per lane (conflict-free thanks to the swizzle) and tcgen05.st's them into the fill's TMEM columns.
        constexpr int NRW = 32 / RPI;                   // LDGSTS per lane for the 32 rows of one chunk
        constexpr int COLS = ROW / 4;                   // TMEM columns per chunk row
        const uint32_t group = (uint32_t)warp >> 2, quarter = (uint32_t)warp & 3u;
        const int piece = lane % CH, sub = lane / CH;
        const uint32_t QB = (uint32_t)P.qb;
        const uint32_t stage0 = L.a0 + (uint32_t)warp * (2u * QB * 32u * ROW);
        // write side: local row lr = i * RPI + sub, piece p sits at p ^ swz(lr); read side: lane reads local row `lane`
        const uint32_t w_off_even = (uint32_t)(sub * ROW + ((ROW == 128 ? (piece ^ sub) : (piece ^ ((sub >> 1) & 3))) << 4));
        const uint32_t w_off_odd = (uint32_t)(sub * ROW + ((ROW == 128 ? (piece ^ (sub + 4)) : (piece ^ ((sub >> 1) & 3))) << 4));
        const uint32_t r_swz = ROW == 128 ? (uint32_t)(lane & 7) : (uint32_t)((lane >> 1) & 3);
        const uint64_t zero_src = (uint64_t)g_zero_rows + (uint32_t)(piece * 16);
        const uint32_t G = (uint32_t)P.groups;
        const uint32_t lane_taddr = tmem_base + ((quarter * 32u) << 16) + (uint32_t)P.a_col0;
        uint32_t my_next = group, c0 = 0, slot = group, phase = 0, witer = 0, nbatch = 0;
        // One batch of copies is always in flight across batch, fill and work-item boundaries: batch b+1 is issued before
        // batch b is retired (staging -> registers -> TMEM; the first batch of a fill first waits for the fill slot, the
        // last one hands the slot to the MMA warp).
        uint32_t pend_sb = 0, pend_bc = 0, pend_pos = 0, pend_slot = 0, pend_parity = 0;
        bool pend_last = false;
        auto retire = [&]() {
            if (pend_pos == 0) {
                mbar_wait(L.empty(pend_slot), pend_parity);    // the MMAs that read this fill slot last time have retired
                tc_fence_after();
            }
            for (uint32_t c = 0; c < pend_bc; ++c) {
                const uint32_t rowaddr = stage0 + (pend_sb * QB + c) * (32u * ROW) + (uint32_t)lane * ROW;
                const uint32_t taddr = lane_taddr + (pend_slot * Q + pend_pos + c) * COLS;
#pragma unroll
                for (int h = 0; h < ROW / 64; ++h) {           // 64 B = 16 columns per store
                    const uint4 v0 = ld_shared_u4(rowaddr + (((uint32_t)(4 * h + 0) ^ r_swz) << 4));
                    const uint4 v1 = ld_shared_u4(rowaddr + (((uint32_t)(4 * h + 1) ^ r_swz) << 4));
                    const uint4 v2 = ld_shared_u4(rowaddr + (((uint32_t)(4 * h + 2) ^ r_swz) << 4));
                    const uint4 v3 = ld_shared_u4(rowaddr + (((uint32_t)(4 * h + 3) ^ r_swz) << 4));
                    tmem_st16(taddr + 16 * h, v0, v1, v2, v3);
                }
            }
            if (pend_last) {
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (elect_one()) mbar_arrive(L.full(pend_slot));
                __syncwarp();
            }
            pend_bc = 0;
        };
        for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++witer) {
            const int tile = w / P.splits;
            const int64_t row0 = (int64_t)tile * BM + quarter * 32;
            const uint32_t buf = witer % NDESC;
            mbar_wait_sleep(L.wfull(buf), (witer / NDESC) & 1u, (uint32_t)P.sleep_ns);
            const uint32_t n = ld_shared_u32(L.count(buf));
            const uint32_t nfill = (n + Q - 1) / Q;
            // gather one batch of chunks [jb, jb + bc) of the item's list into staging buffer sb
            auto issue_batch = [&](uint32_t jb, uint32_t bc, uint32_t sb) {
                for (uint32_t c = 0; c < bc; ++c) {
                    const uint32_t e = ld_shared_u32(L.list(buf, (int)(jb + c)));
                    const uint32_t s = e >> 8, k = (e >> 3) & 31u, kb = e & 7u;
                    const uint2 sa = ld_shared_u2(L.segtab + 32 * s);
                    const uint2 sbv = ld_shared_u2(L.segtab + 32 * s + 16);
                    const uint64_t src0 = ((uint64_t)sa.y << 32 | sa.x) + (uint32_t)(piece * 16) + kb * ROW;
                    const uint32_t seg_stride = sbv.x, seg_idx = ld_shared_u32(L.segtab + 32 * s + 24);
                    const uint32_t dst0 = stage0 + (sb * QB + c) * (32u * ROW);
                    if (seg_idx != 0xffffffffu) {
                        const uint32_t ia = L.idx(buf, (int)(seg_idx + k), (int)(quarter * 32) + sub);
                        int r[NRW];
#pragma unroll
                        for (int i = 0; i < NRW; ++i) r[i] = ld_shared_i32(ia + (uint32_t)(i * RPI * 4));
#pragma unroll
                        for (int i = 0; i < NRW; ++i) {
                            const uint32_t dst = dst0 + (uint32_t)(i * RPI * ROW) + ((i & 1) ? w_off_odd : w_off_even);
                            const uint64_t src = r[i] >= 0 ? src0 + (uint64_t)(uint32_t)r[i] * seg_stride
                                                           : zero_src + kb * ROW + ((((uint32_t)tile * 37u + quarter * 61u + (uint32_t)sub * 16u + i) & 255u) << 10);
                            cp_async16_cg(dst, reinterpret_cast<const void*>(src), 16u);
                        }
                    } else {   // identity segment: row = tile row
#pragma unroll
                        for (int i = 0; i < NRW; ++i) {
                            const int64_t row = row0 + i * RPI + sub;
                            const uint32_t dst = dst0 + (uint32_t)(i * RPI * ROW) + ((i & 1) ? w_off_odd : w_off_even);
                            const uint64_t src = src0 + (uint64_t)(uint32_t)(row < d.n_out ? row : 0) * seg_stride;
                            cp_async16_cg(dst, reinterpret_cast<const void*>(src), row < d.n_out ? 16u : 0u);
                        }
                    }
                }
                cp_async_commit();
            };
            while (my_next < c0 + nfill) {
                const uint32_t j0 = (my_next - c0) * Q;
                const uint32_t cnt = min(Q, n - j0);
                for (uint32_t b0 = 0; b0 < cnt; b0 += QB) {
                    const uint32_t bc = min(QB, cnt - b0);
                    __syncwarp();                              // every lane is done reading the buffer this batch overwrites
                    issue_batch(j0 + b0, bc, nbatch & 1u);
                    if (pend_bc) {                             // the batch issued one step earlier has landed by now (mostly)
                        cp_async_wait<1>();
                        __syncwarp();
                        retire();
                    }
                    pend_sb = nbatch & 1u, pend_bc = bc, pend_pos = b0, pend_slot = slot, pend_parity = phase ^ 1u;
                    pend_last = b0 + QB >= cnt;
                    ++nbatch;
                }
                my_next += G;
                slot += G;
                if (slot >= S) slot -= S, phase ^= 1u;
            }
            c0 += nfill;
            mbar_arrive(L.wempty(buf));       // the item's rulebook rows / chunk list are no longer needed once its copies are issued
        }
        if (pend_bc) {
            cp_async_wait<0>();
            __syncwarp();
            retire();
        }
    } else if (!P.ts && warp < P.q * P.groups) {'''


def main():
    s = open(SRC).read()
    s = sub(s, '#include "tl_common.cuh"', '#include "../../../treelearn_b200/csrc/tl_common.cuh"')
    # -- launch parameters
    s = sub(s, '    int zero_row;      // 1:', '    int ts, qb, a_col0, a_substages;   // TS form: on/off, chunks per gather batch, first TMEM column of the A ring, A staging size in chunk stages\n    int zero_row;      // 1:')
    # -- helpers
    s = sub(s, '__device__ __forceinline__ void umma_commit(uint32_t bar) {', TS_HELPERS + '__device__ __forceinline__ void umma_commit(uint32_t bar) {')
    # -- shared-memory carve-up: the A region is sized separately (private staging in TS form)
    s = sub(s, '__device__ __forceinline__ Layout carve(uint32_t base, int n, int substages, int row_bytes) {',
            '__device__ __forceinline__ Layout carve(uint32_t base, int n, int substages, int row_bytes, int a_substages) {')
    s = sub(s, '    L.b0 = base + substages * L.a_stage_bytes;', '    L.b0 = base + a_substages * L.a_stage_bytes;')
    s = sub(s, 'static inline size_t smem_bytes(int n, int substages, int row_bytes) {\n    return 1024 + (size_t)substages * ((size_t)BM * row_bytes + (size_t)n * row_bytes)',
            'static inline size_t smem_bytes(int n, int substages, int row_bytes, int a_substages) {\n    return 1024 + (size_t)a_substages * BM * row_bytes + (size_t)substages * n * row_bytes')
    s = sub(s, '    const Layout L = carve(base, N, P.stages * P.q, ROW);', '    const Layout L = carve(base, N, P.stages * P.q, ROW, P.a_substages);')
    # -- barrier counts
    s = sub(s, '            mbar_init(L.full(s), 32 * P.q + 1);', '            mbar_init(L.full(s), P.ts ? 4 + 1 : 32 * P.q + 1);   // TS: one elected lane per warp of the group')
    s = sub(s, '            mbar_init(L.wempty(b), 32 * P.q * P.groups + kEpilogueThreads + 2);',
            '            mbar_init(L.wempty(b), 32 * (P.ts ? 4 : P.q) * P.groups + kEpilogueThreads + 2);')
    # -- producers
    s = sub(s, '    } else if (warp < P.q * P.groups) {', TS_PRODUCER)
    # -- MMA issue: TS form reads A from the fill's TMEM columns
    s = sub(s, '                        if (!(P.debug & 1)) {\n                            constexpr uint32_t A_STEP',
            '''                        if (P.ts && !(P.debug & 1)) {
                            constexpr uint32_t COLS = ROW / 4;
                            const uint32_t a_t = tmem_base + (uint32_t)P.a_col0 + slot * Q * COLS;
                            const uint64_t bdesc = bdesc0 + (uint64_t)(slot * Q * b_step);
                            for (uint32_t qi = 0; qi < cnt; ++qi) {
#pragma unroll
                                for (int kk = 0; kk < KSTEPS; ++kk)
                                    umma_f16_ts(tmem_d, a_t + qi * COLS + 8u * kk, bdesc + (uint64_t)(qi * b_step) + (uint64_t)(kk * 2), idesc,
                                                (j0 == 0 && qi == 0 && kk == 0) ? 0u : 1u);
                            }
                        } else if (!(P.debug & 1)) {
                            constexpr uint32_t A_STEP''')
    # -- host: parameters of the TS form
    s = sub(s, '    const size_t smem = tc::smem_bytes(n, stages * q, row_bytes);',
            '''    P.ts = 0, P.qb = 0, P.a_col0 = 0, P.a_substages = stages * q;
    if (half && n <= env_int("TL_TS_MAX_N", 64) && env_int("TL_TS", 1)) {
        // fills of Q chunks in S = 3 TMEM slots behind the two accumulators; A staging = 12 warps x 2 buffers x QB chunks x 32 rows
        const int cols = row_bytes / 4, S = 3;
        int Q = (512 - 2 * P.buf_cols) / cols / S;
        if (Q > env_int("TL_TS_Q", 8)) Q = env_int("TL_TS_Q", 8);
        const int QB = env_int("TL_TS_QB", row_bytes == 64 ? 2 : 1);
        const int a_sub = tc::kProducerWarps * 2 * QB / 4;        // in 128-row chunk stages
        const int budget = env_int("TL_TS_SMEM_KB", 208) * 1024;
        while (Q > 1 && tc::smem_bytes(n, S * Q, row_bytes, a_sub) > (size_t)budget) --Q;
        if (Q >= 1 && tc::smem_bytes(n, S * Q, row_bytes, a_sub) <= 227 * 1024) {
            P.ts = 1, P.qb = QB, P.a_col0 = 2 * P.buf_cols, P.a_substages = a_sub;
            P.acc_ways = 1, P.buf_cols = P.acc_cols, P.a_col0 = 2 * P.buf_cols;
            P.tmem_cols = 512;
            P.q = Q, P.stages = S, P.groups = 3;
            stages = S, q = Q;
        }
    }
    const size_t smem = tc::smem_bytes(n, stages * q, row_bytes, P.a_substages);''')
    s = sub(s, '    while (tc::smem_bytes(n, stages * q, row_bytes) > 227 * 1024 && stages > 2) --stages;',
            '    while (tc::smem_bytes(n, stages * q, row_bytes, stages * q) > 227 * 1024 && stages > 2) --stages;')
    # -- the experimental library: own entry point + the two helpers tl_common.cuh expects, no SIMT fallback
    s = sub(s, 'int conv_fwd_simt(const tl_conv_desc& d, cudaStream_t stream);',
            '''static char g_err[512];
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int) {}
static int conv_fwd_simt(const tl_conv_desc&, cudaStream_t) {
    set_error("experimental TS library: shape needs the SIMT path");
    return TL_ERR_UNSUPPORTED;
}''')
    s = sub(s, '#include <stdlib.h>\n', '#include <stdarg.h>\n#include <stdlib.h>\n')
    s = sub(s, 'extern "C" int tl_debug_copy_trace(void* host, size_t bytes) {', 'extern "C" int tl_debug_copy_trace_ts(void* host, size_t bytes) {')
    s += '''
// mode 1 = tf32 (always the shared-memory form), 2 = fp16 (TS form where eligible)
extern "C" __attribute__((visibility("default"))) int tl_conv_fwd_ts(const tl_conv_desc* d, int mode, void* stream) {
    return tl::conv_fwd_tc(*d, (cudaStream_t)stream, mode == 2);
}
extern "C" __attribute__((visibility("default"))) const char* tl_last_error_ts(void) { return tl::g_err; }
'''
    os.makedirs(OUT_DIR, exist_ok=True)
    out = os.path.join(OUT_DIR, 'tl_conv_tc_ts.cu')
    open(out, 'w').write(s)
    so = os.path.join(OUT_DIR, 'libtl_conv_ts.so')
    cmd = ['nvcc', '-O3', '-std=c++17', '-lineinfo', '-gencode', 'arch=compute_100a,code=sm_100a', '-Xcompiler', '-fPIC',
           '-Xcompiler', '-fvisibility=hidden', '--expt-relaxed-constexpr', '-Xptxas', '-v', '-shared', '-o', so, out, '-lcudart', '-lcuda']
    print(' '.join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    open(os.path.join(OUT_DIR, 'ptxas.log'), 'w').write(log)
    print('\n'.join(l for l in log.splitlines() if 'error' in l.lower() or 'warning' in l.lower() or 'k_conv_tc' in l or 'registers' in l)[:4000])
    sys.exit(res.returncode)


if __name__ == '__main__':
    main()
