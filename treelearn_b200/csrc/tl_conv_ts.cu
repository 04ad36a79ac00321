// tcgen05 sparse convolution with the A operand in TENSOR MEMORY (TS-form tcgen05.mma), sm_100a only.
// Modes TL_MODE_F16 (fp16 operands) and TL_MODE_F16X2 (two-term fp16 split of both operands: hi + lo, three MMAs per
// K step = fp32-equivalent products), fp32 accumulation in both.
//
// Why (profiles/r01_conv_tc_history.md, profiles/r02_ts_probe.txt): with the gathered A tile in shared memory every
// M128 x N x K16 MMA streams 4 KB of A through the tensor proxy at 64 B/clk = 64 cycles, whatever N is; the C_out = 32 / 64
// layers (80 % of the bytes of the U-Net) ran the tensor pipe at 25-50 % of its rate and the kernel at 14.5 % of the HBM
// roofline.  Here the gathered rows never touch shared memory:
//   * transfer warps (16; warp % 4 = TMEM lane quarter) load the neighbour rows straight from global / L1 into
//     registers -- a quad of lanes reads one contiguous 64 B row piece, 8 rows per LDG.128, absent neighbours are
//     predicated off (no traffic at all) -- and write them into a TMEM ring with tcgen05.st.16x256b (the fragment layout
//     of that shape is exactly "4 lanes per row, 8 B each");
//   * the MMA warp issues tcgen05.mma [d_tmem], [a_tmem], b_desc: A from TMEM, B (weights) from shared memory, where the
//     weights of the whole layer stay resident when they fit (C = 32: 54 KB) and stream through a ring otherwise;
//   * a ring slot ("fill") is Q chunks (a chunk = 128 rows x 32 channels of one (segment, offset)) handed over with ONE
//     mbarrier round trip and ONE tcgen05.commit, so the fixed cost of the single-thread MMA loop is paid once per Q chunks.
// The K order inside a 32-channel chunk is permuted (K step kk, position 4q+e <-> channel 8q + 4kk + e) because that is
// what the 16x256b fragment gives for a contiguous 16 B read per lane; the weights are packed with the same permutation
// (sparse.pack_weight_ts), so the product is unchanged.
// Scheduler, rulebook staging, split-K and the coalescing epilogue follow tl_conv_tc.cu.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "tl_common.cuh"
#include "tl_tc_ptx.cuh"

namespace tl {
namespace ts {

using namespace tl::tc;

constexpr int kEpilogueThreads = 128;
constexpr int MAX_SLOTS = 8;
constexpr int MAX_CHUNKS = 448;            // live (segment, offset, k-block) entries per work item
constexpr int IDX_ROWS = 28;               // rulebook rows staged per work item
constexpr int IDX_BUF_BYTES = IDX_ROWS * BM * 4 + MAX_CHUNKS * 4 + 64;
constexpr int NDESC = 3;
constexpr int EPI_BYTES = 4 * 32 * 128;
constexpr int SEGTAB_BYTES = 32 * TL_MAX_SEG;
constexpr int TMEM_COLS = 512;
constexpr int NTW1 = 16, NTW2 = 12;   // transfer warps of the fp16 / split-format kernels

struct Launch {
    int num_tiles, splits, chunks_total;
    int S, Q;                    // A ring: S slots ("fills") of Q chunks
    int acc_bufs, acc_stride;    // accumulator buffers (1 or 2) and their column stride
    int a_col0;                  // first TMEM column of the A ring
    int resident;                // 1: every weight slab of the layer stays in shared memory; 0: weight ring (one stage per ring chunk)
    uint32_t b_bytes;            // shared-memory bytes of the weight region
    int debug;                   // timing experiments: 1 skip MMAs, 2 skip row loads, 4 skip epilogue memory ops, 8 skip weight copies, 16 skip tcgen05.st
    int sleep_ns;
    float* splitk_ws;
    int idx_base[TL_MAX_SEG], idx_owner[TL_MAX_SEG];
    int src_fp32[TL_MAX_SEG];    // the segment's source rows are raw fp32 (converted to the operand format in registers)
};

struct Layout {
    uint32_t b0, idx0, epi0, bars, tmem_slot, segtab;
    __device__ __forceinline__ uint32_t idx(uint32_t buf, int row, int col) const {
        return idx0 + buf * IDX_BUF_BYTES + (uint32_t)(row * BM + col) * 4u;
    }
    __device__ __forceinline__ uint32_t list(uint32_t buf, int j) const {
        return idx0 + buf * IDX_BUF_BYTES + IDX_ROWS * BM * 4 + (uint32_t)j * 4u;
    }
    __device__ __forceinline__ uint32_t count(uint32_t buf) const {
        return idx0 + buf * IDX_BUF_BYTES + IDX_ROWS * BM * 4 + MAX_CHUNKS * 4;
    }
    __device__ __forceinline__ uint32_t full(uint32_t s) const { return bars + 8 * s; }
    __device__ __forceinline__ uint32_t empty(uint32_t s) const { return bars + 8 * (MAX_SLOTS + s); }
    __device__ __forceinline__ uint32_t tfull(uint32_t b) const { return bars + 8 * (2 * MAX_SLOTS + b); }
    __device__ __forceinline__ uint32_t tempty(uint32_t b) const { return bars + 8 * (2 * MAX_SLOTS + 2 + b); }
    __device__ __forceinline__ uint32_t wfull(uint32_t b) const { return bars + 8 * (2 * MAX_SLOTS + 4 + b); }
    __device__ __forceinline__ uint32_t wempty(uint32_t b) const { return bars + 8 * (2 * MAX_SLOTS + 4 + NDESC + b); }
    __device__ __forceinline__ uint32_t wres() const { return bars + 8 * (2 * MAX_SLOTS + 4 + 2 * NDESC); }
};
constexpr int BAR_BYTES = (8 * (2 * MAX_SLOTS + 4 + 2 * NDESC + 1) + 15) & ~15;   // keeps the segment table 16 B aligned

__device__ __forceinline__ Layout carve(uint32_t base, uint32_t b_bytes) {
    Layout L;
    L.b0 = base;
    L.idx0 = L.b0 + b_bytes;
    L.epi0 = L.idx0 + NDESC * IDX_BUF_BYTES;
    L.bars = L.epi0 + EPI_BYTES;
    L.tmem_slot = L.bars + BAR_BYTES;
    L.segtab = L.tmem_slot + 16;
    return L;
}
static inline size_t smem_bytes(size_t b_bytes) {
    return 1024 + b_bytes + NDESC * IDX_BUF_BYTES + EPI_BYTES + BAR_BYTES + 16 + SEGTAB_BYTES + 32;
}

__device__ __forceinline__ uint32_t seg_mask(const tl_conv_seg& sg, int64_t tile) {
    if (!sg.index) return 1u;
    const uint32_t all = sg.n_off >= 32 ? 0xffffffffu : ((1u << sg.n_off) - 1u);
    return (sg.tile_mask ? sg.tile_mask[tile] : 0xffffffffu) & all;
}
__device__ __forceinline__ void st_shared_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_shared_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 ldg_nc_v4(uint64_t addr) {
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(addr));
    return v;
}
// tcgen05.mma with the A operand in tensor memory
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 16 lanes x 256 bit x 2: thread t holds, per 8-column group g, columns 2(t%4), 2(t%4)+1 of lane t/4 (a[g]) and of lane t/4 + 8 (b[g])
__device__ __forceinline__ void tmem_st_16x256b_x2(uint32_t taddr, uint32_t a0x, uint32_t a0y, uint32_t b0x, uint32_t b0y,
                                                   uint32_t a1x, uint32_t a1y, uint32_t b1x, uint32_t b1y) {
    asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(a0x),
                 "r"(a0y), "r"(b0x), "r"(b0y), "r"(a1x), "r"(a1y), "r"(b1x), "r"(b1y)
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// One gathered chunk of a transfer thread: 4 rows (lane quarter rows rr, rr+8, rr+16, rr+24) x 8 channels
template <int NSPLIT>
struct Frag {
    uint4 hi[4];
    uint4 lo[NSPLIT == 2 ? 4 : 1];
};
template <>
struct Frag<0> {};   // placeholder (no second fragment in the split format)

template <int NSPLIT>
__device__ __forceinline__ void load_frag(Frag<NSPLIT>& f, const int (&ix)[4], uint64_t src, uint32_t row_bytes, bool fp32src) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (ix[j] < 0) {
            f.hi[j] = make_uint4(0u, 0u, 0u, 0u);
            if (NSPLIT == 2) f.lo[j] = make_uint4(0u, 0u, 0u, 0u);
            continue;
        }
        const uint64_t p = src + (uint64_t)(uint32_t)ix[j] * row_bytes;
        if (!fp32src) {
            f.hi[j] = ldg_nc_v4(p);
            if (NSPLIT == 2) f.lo[j] = ldg_nc_v4(p + 64);
        } else {   // 8 raw fp32 channels -> operand format
            const uint4 u0 = ldg_nc_v4(p), u1 = ldg_nc_v4(p + 16);
            const float v0 = __uint_as_float(u0.x), v1 = __uint_as_float(u0.y), v2 = __uint_as_float(u0.z), v3 = __uint_as_float(u0.w);
            const float v4 = __uint_as_float(u1.x), v5 = __uint_as_float(u1.y), v6 = __uint_as_float(u1.z), v7 = __uint_as_float(u1.w);
            if (NSPLIT == 1) {
                f.hi[j] = make_uint4(pack_half2(v0, v1), pack_half2(v2, v3), pack_half2(v4, v5), pack_half2(v6, v7));
            } else {
                split_half2(v0, v1, f.hi[j].x, f.lo[j].x);
                split_half2(v2, v3, f.hi[j].y, f.lo[j].y);
                split_half2(v4, v5, f.hi[j].z, f.lo[j].z);
                split_half2(v6, v7, f.hi[j].w, f.lo[j].w);
            }
        }
    }
}
// taddr: TMEM address of the chunk's first column in this warp's lane quarter (lane field = 32 * quarter)
template <int NSPLIT>
__device__ __forceinline__ void store_frag(const Frag<NSPLIT>& f, uint32_t taddr) {
    // K step 0 <- bytes [0, 8) of the lane's 16 B piece, K step 1 <- bytes [8, 16)
    tmem_st_16x256b_x2(taddr, f.hi[0].x, f.hi[0].y, f.hi[1].x, f.hi[1].y, f.hi[0].z, f.hi[0].w, f.hi[1].z, f.hi[1].w);
    tmem_st_16x256b_x2(taddr + (16u << 16), f.hi[2].x, f.hi[2].y, f.hi[3].x, f.hi[3].y, f.hi[2].z, f.hi[2].w, f.hi[3].z, f.hi[3].w);
    if (NSPLIT == 2) {
        tmem_st_16x256b_x2(taddr + 16, f.lo[0].x, f.lo[0].y, f.lo[1].x, f.lo[1].y, f.lo[0].z, f.lo[0].w, f.lo[1].z, f.lo[1].w);
        tmem_st_16x256b_x2(taddr + 16 + (16u << 16), f.lo[2].x, f.lo[2].y, f.lo[3].x, f.lo[3].y, f.lo[2].z, f.lo[2].w, f.lo[3].z,
                           f.lo[3].w);
    }
}

// Warp roles (736 threads, one CTA per SM, persistent over work items = (row tile, K split)):
//   warps  0..15  transfer: warp w = (group w / 4, lane quarter w % 4); group g moves chunks g, g + G, ... of every fill
//   warps 16..19  epilogue (TMEM lane quarter = warp % 4)
//   warp  20      MMA issuer + TMEM alloc/dealloc
//   warp  21      scheduler (rulebook rows + live chunk list of the work item, two items ahead)
//   warp  22      weight loader (TMA bulk copies: everything once when resident, else one slab per ring chunk)
// NTW = transfer warps (a multiple of 4: G = NTW / 4 groups x 4 lane quarters); 16 for fp16 (88 registers per thread),
// 12 for the split format, whose chunk fragment is twice as many registers
template <int NSPLIT, int NTW>
__global__ void __launch_bounds__(32 * (NTW + 7), 1) k_conv_ts(const tl_conv_desc d, const Launch P) {
    constexpr uint32_t CHUNK_COLS = 16 * NSPLIT;       // TMEM columns of one A chunk
    constexpr uint32_t G = NTW / 4;
    constexpr int kFirstEpilogueWarp = NTW;            // NTW % 4 == 0, so epilogue warp e owns TMEM lane quarter e
    constexpr int kMmaWarp = NTW + 4, kSchedWarp = NTW + 5, kWeightWarp = NTW + 6;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int N = d.c_out;
    const Layout L = carve(base, P.b_bytes);
    const uint32_t Q = (uint32_t)P.Q, S = (uint32_t)P.S;
    const uint32_t slab = (uint32_t)N * 64u * NSPLIT;   // weight bytes of one chunk: [C_out][32] fp16 (x hi, lo)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (L.tmem_slot - smem_u32(smem_raw)));
    const int num_work = P.num_tiles * P.splits;
    const int per_split = (P.chunks_total + P.splits - 1) / P.splits;

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < S; ++s) {
            mbar_init(L.full(s), NTW + (P.resident ? 0 : 1));   // one elected lane per transfer warp (+ the weight loader's expect_tx)
            mbar_init(L.empty(s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(L.tfull(b), 1);
            mbar_init(L.tempty(b), kEpilogueThreads);
        }
        for (int b = 0; b < NDESC; ++b) {
            mbar_init(L.wfull(b), 33);
            mbar_init(L.wempty(b), NTW + kEpilogueThreads + 2);   // transfer warps (elected lane), epilogue threads, MMA + weight warps
        }
        mbar_init(L.wres(), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < d.n_seg) {
        const tl_conv_seg& sg = d.seg[threadIdx.x];
        const uint32_t t = L.segtab + 32 * threadIdx.x;
        const uint64_t src = (uint64_t)sg.src;
        const uint32_t fp32src = (uint32_t)P.src_fp32[threadIdx.x];
        // bytes per source row / per 32-channel block of it
        const uint32_t row_bytes = (uint32_t)sg.src_stride * (fp32src ? 4u : 2u * NSPLIT);
        st_shared_v4(t, (uint32_t)src, (uint32_t)(src >> 32), row_bytes, fp32src);
        st_shared_v4(t + 16, sg.index ? (uint32_t)P.idx_base[threadIdx.x] : 0xffffffffu, 0u, 0u, 0u);
    }
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(L.tmem_slot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == kSchedWarp) {
        // ===================== scheduler: work-item descriptors ====================================
        uint32_t witer = 0;
        for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++witer) {
            const uint32_t buf = witer % NDESC;
            if (witer >= NDESC) mbar_wait_sleep(L.wempty(buf), ((witer / NDESC) - 1u) & 1u, (uint32_t)P.sleep_ns);
            const int tile = w / P.splits, split = w - tile * P.splits;
            const int lo = split * per_split, hi = min(lo + per_split, P.chunks_total);
            int ord = 0, pos = 0;
            for (int s = 0; s < d.n_seg; ++s) {
                const tl_conv_seg& sg = d.seg[s];
                const int kblocks = sg.c_in / 32;
                if (sg.index && P.idx_owner[s]) {
                    const int32_t* ip = sg.index + (int64_t)tile * BM + 4 * lane;
                    for (int k = 0; k < sg.n_off; ++k)
                        cp_async16(L.idx(buf, P.idx_base[s] + k, 4 * lane), ip + (int64_t)k * sg.index_stride, 16u);
                }
                const uint32_t mask = seg_mask(sg, tile);
                const int o = ord + lane * kblocks;
                int kb_lo = 0, cnt = 0;
                if (lane < sg.n_off && ((mask >> lane) & 1u)) {
                    kb_lo = max(lo - o, 0);
                    cnt = max(min(hi - o, kblocks) - kb_lo, 0);
                }
                int incl = cnt;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, incl, off);
                    if (lane >= off) incl += v;
                }
                const int mypos = pos + incl - cnt;
                for (int t = 0; t < cnt; ++t)
                    if (mypos + t < MAX_CHUNKS)   // entry: weight slab ordinal | segment | offset | k-block
                        st_shared_u32(L.list(buf, mypos + t), ((uint32_t)(o + kb_lo + t) << 16) | (uint32_t)((s << 8) | (lane << 3) | (kb_lo + t)));
                pos += __shfl_sync(0xffffffffu, incl, 31);
                ord += sg.n_off * kblocks;
            }
            if (lane == 0) st_shared_u32(L.count(buf), (uint32_t)min(pos, MAX_CHUNKS));
            __syncwarp();
            cp_async_mbar_arrive_noinc(L.wfull(buf));
            if (lane == 0) mbar_arrive(L.wfull(buf));
        }
    } else if (warp < NTW) {
        // ===================== transfer: global rows -> registers -> TMEM A ring ====================
        const uint32_t quarter = (uint32_t)warp & 3u, grp = (uint32_t)warp >> 2;
        const int q = lane & 3, rr = lane >> 2;
        const uint32_t lane_field = (quarter * 32u) << 16;
        uint32_t fill = 0;                      // ordinal of the fill in this CTA's stream: slot = fill % S
        uint32_t slot = 0, phase = 0;
        uint32_t witer = 0;
        uint32_t prev_s = 0xffffffffu, seg_row_bytes = 0, seg_idx = 0xffffffffu;
        uint64_t seg_src = 0;
        bool seg_fp32 = false;
        for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++witer) {
            const int tile = w / P.splits;
            const uint32_t buf = witer % NDESC;
            mbar_wait_sleep(L.wfull(buf), (witer / NDESC) & 1u, (uint32_t)P.sleep_ns);
            const uint32_t n = ld_shared_u32(L.count(buf));
            const int trow = tile * BM + (int)quarter * 32 + rr;          // identity segments: my first row
            for (uint32_t j0 = 0; j0 < n; j0 += Q, ++fill) {
                const uint32_t cnt = min(Q, n - j0);
                Frag<NSPLIT> fa;
                Frag<(NSPLIT == 1 ? 1 : 0)> fb;
                uint32_t i = grp;
                // software pipeline over this warp's chunks of the fill: the loads of chunk i + G are in flight while
                // chunk i is written to TMEM; the slot itself is only needed by the first store
                auto fetch = [&](Frag<NSPLIT>& f, uint32_t ci) {
                    const uint32_t e = ld_shared_u32(L.list(buf, (int)(j0 + ci)));
                    const uint32_t s = (e >> 8) & 0xffu, k = (e >> 3) & 31u, kb = e & 7u;
                    if (s != prev_s) {
                        prev_s = s;
                        uint32_t a, b, c, f32;
                        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(f32) : "r"(L.segtab + 32 * s));
                        seg_src = ((uint64_t)b << 32) | a;
                        seg_row_bytes = c;
                        seg_fp32 = f32 != 0u;
                        seg_idx = ld_shared_u32(L.segtab + 32 * s + 16);
                    }
                    int ix[4];
                    if (seg_idx != 0xffffffffu) {
                        const uint32_t ia = L.idx(buf, (int)(seg_idx + k), (int)quarter * 32 + rr);
#pragma unroll
                        for (int j = 0; j < 4; ++j) ix[j] = ld_shared_i32(ia + 32u * j);
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) ix[j] = (trow + 8 * j < d.n_out) ? trow + 8 * j : -1;
                    }
                    if (P.debug & 2) ix[0] = ix[1] = ix[2] = ix[3] = -1;
                    const uint64_t src = seg_src + (seg_fp32 ? (uint64_t)(kb * 128u + q * 32u) : (uint64_t)(kb * 64u * NSPLIT + q * 16u));
                    load_frag<NSPLIT>(f, ix, src, seg_row_bytes, seg_fp32);
                };
                if (i < cnt) fetch(fa, i);
                mbar_wait(L.empty(slot), phase ^ 1u);
                tc_fence_after();
                const uint32_t ring = tmem_base + lane_field + (uint32_t)P.a_col0 + slot * Q * CHUNK_COLS;
                if constexpr (NSPLIT == 1) {
                    while (i < cnt) {
                        uint32_t nx = i + G;
                        if (nx < cnt) fetch(fb, nx);
                        if (!(P.debug & 16)) store_frag<NSPLIT>(fa, ring + i * CHUNK_COLS);
                        i = nx;
                        if (i >= cnt) break;
                        nx = i + G;
                        if (nx < cnt) fetch(fa, nx);
                        if (!(P.debug & 16)) store_frag<NSPLIT>(fb, ring + i * CHUNK_COLS);
                        i = nx;
                    }
                } else {   // split format: one fragment (32 registers) at a time
                    while (i < cnt) {
                        if (!(P.debug & 16)) store_frag<NSPLIT>(fa, ring + i * CHUNK_COLS);
                        i += G;
                        if (i < cnt) fetch(fa, i);
                    }
                }
                tmem_wait_st();
                tc_fence_before();
                if (elect_one()) mbar_arrive(L.full(slot));
                __syncwarp();
                if (++slot == S) slot = 0, phase ^= 1u;
            }
            if (elect_one()) mbar_arrive(L.wempty(buf));
            __syncwarp();
        }
    } else if (warp == kWeightWarp) {
        // ===================== weight loader ========================================================
        // weights are packed [n_off][c_in/32][(hi, lo)][C_out][32] with the K permutation and the SWIZZLE_64B image applied
        // (sparse.pack_weight_ts): the slabs of a segment are contiguous, a slab lands in its stage as-is
        if (P.resident) {
            if (elect_one()) {
                uint32_t total = 0;
                for (int s = 0; s < d.n_seg; ++s) total += (uint32_t)(d.seg[s].n_off * (d.seg[s].c_in / 32)) * slab;
                if (!(P.debug & 8)) {
                    mbar_arrive_expect_tx(L.wres(), total);
                    uint32_t dst = L.b0;
                    for (int s = 0; s < d.n_seg; ++s) {
                        const uint32_t bytes = (uint32_t)(d.seg[s].n_off * (d.seg[s].c_in / 32)) * slab;
                        const char* src = reinterpret_cast<const char*>(d.seg[s].weight);
                        for (uint32_t off = 0; off < bytes; off += 32768u) {
                            const uint32_t nb = min(32768u, bytes - off);
                            bulk_g2s(dst + off, src + off, nb, L.wres());
                        }
                        dst += bytes;
                    }
                } else {
                    mbar_arrive(L.wres());
                }
            }
            __syncwarp();
            // descriptors are still consumed (and released) per work item so that the scheduler's ring keeps turning
            uint32_t witer = 0;
            for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++witer) {
                const uint32_t buf = witer % NDESC;
                mbar_wait_sleep(L.wfull(buf), (witer / NDESC) & 1u, (uint32_t)P.sleep_ns);
                if (elect_one()) mbar_arrive(L.wempty(buf));
                __syncwarp();
            }
        } else {
            uint64_t wbase[TL_MAX_SEG];
            uint32_t wfirst[TL_MAX_SEG];   // slab ordinal of the segment's first slab
            uint32_t acc = 0;
#pragma unroll
            for (int s = 0; s < TL_MAX_SEG; ++s) {
                wbase[s] = s < d.n_seg ? (uint64_t)d.seg[s].weight : 0;
                wfirst[s] = acc;
                if (s < d.n_seg) acc += (uint32_t)(d.seg[s].n_off * (d.seg[s].c_in / 32));
            }
            uint32_t slot = 0, phase = 0, witer = 0;
            for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++witer) {
                const uint32_t buf = witer % NDESC;
                mbar_wait_sleep(L.wfull(buf), (witer / NDESC) & 1u, (uint32_t)P.sleep_ns);
                const uint32_t n = ld_shared_u32(L.count(buf));
                for (uint32_t j0 = 0; j0 < n; j0 += Q) {
                    const uint32_t cnt = min(Q, n - j0);
                    uint64_t wsrc = 0;
                    if ((uint32_t)lane < cnt) {
                        const uint32_t e = ld_shared_u32(L.list(buf, (int)(j0 + lane)));
                        const uint32_t s = (e >> 8) & 0xffu, ordn = e >> 16;
                        const uint64_t wb = s == 0 ? wbase[0] : (s == 1 ? wbase[1] : wbase[2]);
                        const uint32_t wf = s == 0 ? wfirst[0] : (s == 1 ? wfirst[1] : wfirst[2]);
                        wsrc = wb + (uint64_t)(ordn - wf) * slab;
                    }
                    mbar_wait(L.empty(slot), phase ^ 1u);
                    if (P.debug & 8) {
                        if (lane == 0) mbar_arrive(L.full(slot));
                    } else {
                        if (lane == 0) mbar_arrive_expect_tx(L.full(slot), cnt * slab);
                        __syncwarp();
                        if ((uint32_t)lane < cnt)
                            bulk_g2s(L.b0 + (slot * Q + lane) * slab, reinterpret_cast<const void*>(wsrc), slab, L.full(slot));
                    }
                    __syncwarp();
                    if (++slot == S) slot = 0, phase ^= 1u;
                }
                if (elect_one()) mbar_arrive(L.wempty(buf));
                __syncwarp();
            }
        }
    } else if (warp >= kFirstEpilogueWarp && warp < kFirstEpilogueWarp + 4) {
        // ===================== epilogue: TMEM -> registers -> (smem transpose) -> global ============
        const int ew = warp - kFirstEpilogueWarp;
        const uint32_t epi = L.epi0 + (uint32_t)ew * (32 * 128);
        const int cc = lane & 7, rsub = lane >> 3;
        uint32_t titer = 0;
        for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++titer) {
            const int tile = w / P.splits;
            const uint32_t abuf = P.acc_bufs == 2 ? (titer & 1u) : 0u;
            const uint32_t aphase = P.acc_bufs == 2 ? ((titer >> 1) & 1u) : (titer & 1u);
            const uint32_t dbuf = titer % NDESC;
            mbar_wait_sleep(L.wfull(dbuf), (titer / NDESC) & 1u, (uint32_t)P.sleep_ns);
            const bool any = ld_shared_u32(L.count(dbuf)) != 0u;
            mbar_arrive(L.wempty(dbuf));
            const int64_t wrow0 = (int64_t)tile * BM + ew * 32;
            const bool has_res = d.residual != nullptr && P.splits == 1 && !(P.debug & 4);
            float4 res[8];
            auto fetch_residual = [&](int c0) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int64_t grow = wrow0 + i * 4 + rsub;
                    res[i] = (has_res && grow < d.n_out) ? __ldg(reinterpret_cast<const float4*>(d.residual + grow * N + c0 + 4 * cc))
                                                         : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            };
            fetch_residual(0);
            mbar_wait_sleep(L.tfull(abuf), aphase, (uint32_t)P.sleep_ns);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + abuf * (uint32_t)P.acc_stride;
            if (any || P.splits == 1) {
                for (int c0 = 0; c0 < N; c0 += 32) {
                    if (c0 > 0) fetch_residual(c0);
                    uint32_t acc[32];
                    if (any) {
                        tmem_ld32(taddr + c0, acc);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) acc[j] = 0u;
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        st_shared_v4(epi + lane * 128 + ((j ^ (lane & 7)) << 4), acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
                    __syncwarp();
                    const int col = c0 + 4 * cc;
                    float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), t1 = s1, s2 = s1, t2 = s1;
                    if (d.out_act1) s1 = __ldg(reinterpret_cast<const float4*>(d.scale1 + col)), t1 = __ldg(reinterpret_cast<const float4*>(d.shift1 + col));
                    if (d.out_act2) s2 = __ldg(reinterpret_cast<const float4*>(d.scale2 + col)), t2 = __ldg(reinterpret_cast<const float4*>(d.shift2 + col));
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int r = i * 4 + rsub;
                        float4 v = ld_shared_f4(epi + r * 128 + ((cc ^ (r & 7)) << 4));
                        const int64_t grow = wrow0 + r;
                        if (grow >= d.n_out || (P.debug & 4)) continue;
                        const int64_t o = grow * N + col;
                        if (P.splits > 1) {
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(P.splitk_ws + o), "f"(v.x), "f"(v.y),
                                         "f"(v.z), "f"(v.w)
                                         : "memory");
                            continue;
                        }
                        v.x += res[i].x, v.y += res[i].y, v.z += res[i].z, v.w += res[i].w;
                        if (d.out_raw) *reinterpret_cast<float4*>(d.out_raw + o) = v;
                        if (d.out_act1)
                            store_act4<(NSPLIT == 2 ? FMT_F16X2 : FMT_F16)>(d.out_act1, grow, N, col, fmaxf(fmaf(v.x, s1.x, t1.x), 0.f), fmaxf(fmaf(v.y, s1.y, t1.y), 0.f),
                                               fmaxf(fmaf(v.z, s1.z, t1.z), 0.f), fmaxf(fmaf(v.w, s1.w, t1.w), 0.f));
                        if (d.out_act2)
                            store_act4<(NSPLIT == 2 ? FMT_F16X2 : FMT_F16)>(d.out_act2, grow, N, col, fmaxf(fmaf(v.x, s2.x, t2.x), 0.f), fmaxf(fmaf(v.y, s2.y, t2.y), 0.f),
                                               fmaxf(fmaf(v.z, s2.z, t2.z), 0.f), fmaxf(fmaf(v.w, s2.w, t2.w), 0.f));
                    }
                    __syncwarp();
                }
            }
            tc_fence_before();
            mbar_arrive(L.tempty(abuf));
        }
    } else if (warp == kMmaWarp) {
        // ===================== MMA issuer (one elected thread) ======================================
        const uint32_t idesc = make_idesc(N, true);
        const uint64_t bdesc0 = make_smem_desc(L.b0, 64);
        const uint32_t slab16 = slab >> 4, half16 = ((uint32_t)N * 64u) >> 4;   // descriptor address units (16 B)
        uint32_t slot = 0, phase = 0, titer = 0;
        if (P.resident) mbar_wait(L.wres(), 0u);
        for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++titer) {
            const uint32_t abuf = P.acc_bufs == 2 ? (titer & 1u) : 0u;
            const uint32_t aphase = P.acc_bufs == 2 ? ((titer >> 1) & 1u) : (titer & 1u);
            const uint32_t dbuf = titer % NDESC;
            mbar_wait_sleep(L.wfull(dbuf), (titer / NDESC) & 1u, (uint32_t)P.sleep_ns);
            const uint32_t n = ld_shared_u32(L.count(dbuf));
            mbar_wait_sleep(L.tempty(abuf), aphase ^ 1u, (uint32_t)P.sleep_ns);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + abuf * (uint32_t)P.acc_stride;
            for (uint32_t j0 = 0; j0 < n; j0 += Q) {
                const uint32_t cnt = min(Q, n - j0);
                // weight slab of chunk `lane` of this fill (resident: by slab ordinal; ring: by stage)
                uint32_t my_b = (slot * Q + (uint32_t)lane) * slab16;
                if (P.resident && (uint32_t)lane < cnt) my_b = (ld_shared_u32(L.list(dbuf, (int)(j0 + lane))) >> 16) * slab16;
                mbar_wait(L.full(slot), phase);
                tc_fence_after();
                const uint32_t a0 = tmem_base + (uint32_t)P.a_col0 + slot * Q * CHUNK_COLS;
                for (uint32_t qi = 0; qi < cnt; ++qi) {
                    const uint32_t boff = __shfl_sync(0xffffffffu, my_b, (int)qi);
                    if (elect_one() && !(P.debug & 1)) {
                        const uint64_t bd = bdesc0 + (uint64_t)boff;
                        const uint32_t a = a0 + qi * CHUNK_COLS;
                        const uint32_t first = (j0 == 0 && qi == 0) ? 0u : 1u;
#pragma unroll
                        for (uint32_t kk = 0; kk < 2; ++kk) {
                            umma_f16_ts(tmem_d, a + 8 * kk, bd + 2 * kk, idesc, kk ? 1u : first);
                            if (NSPLIT == 2) {
                                umma_f16_ts(tmem_d, a + 8 * kk, bd + half16 + 2 * kk, idesc, 1u);        // hi x lo
                                umma_f16_ts(tmem_d, a + 16 + 8 * kk, bd + 2 * kk, idesc, 1u);            // lo x hi
                            }
                        }
                    }
                    __syncwarp();
                }
                if (elect_one()) umma_commit(L.empty(slot));
                __syncwarp();
                if (++slot == S) slot = 0, phase ^= 1u;
            }
            if (elect_one()) {
                mbar_arrive(L.wempty(dbuf));
                if (n == 0) mbar_arrive(L.tfull(abuf));
                else umma_commit(L.tfull(abuf));
            }
            __syncwarp();
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// split-K second pass: v = ws (+ residual) -> raw / act outputs
template <int NSPLIT>
__global__ void k_splitk_epilogue(const tl_conv_desc d, const float* __restrict__ ws) {
    const int64_t total = (int64_t)d.n_out * d.c_out;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int col = (int)(e % d.c_out);
        float v = ws[e];
        if (d.residual) v += __ldg(d.residual + e);
        if (d.out_raw) d.out_raw[e] = v;
        if (d.out_act1) store_act1<(NSPLIT == 2 ? FMT_F16X2 : FMT_F16)>(d.out_act1, e, d.c_out, fmaxf(fmaf(v, __ldg(d.scale1 + col), __ldg(d.shift1 + col)), 0.f));
        if (d.out_act2) store_act1<(NSPLIT == 2 ? FMT_F16X2 : FMT_F16)>(d.out_act2, e, d.c_out, fmaxf(fmaf(v, __ldg(d.scale2 + col), __ldg(d.shift2 + col)), 0.f));
    }
}

}  // namespace ts

static int env_int_ts(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

bool conv_ts_eligible(const tl_conv_desc& d) {
    if (d.c_out % 32 != 0 || d.c_out > 256) return false;
    int idx_rows = 0, chunks = 0;
    for (int s = 0; s < d.n_seg; ++s) {
        const tl_conv_seg& g = d.seg[s];
        if (g.c_in % 32 != 0 || g.c_in > 256 || g.src_stride % 8 != 0) return false;
        if (g.index && g.index_stride % 4 != 0) return false;
        chunks += g.n_off * (g.c_in / 32);
        if (g.index) {
            bool alias = false;
            for (int t = 0; t < s; ++t)
                if (d.seg[t].index == g.index && d.seg[t].index_stride == g.index_stride && d.seg[t].n_off == g.n_off) alias = true;
            if (!alias) idx_rows += g.n_off;
        }
    }
    return idx_rows <= ts::IDX_ROWS && chunks <= ts::MAX_CHUNKS;
}

// nsplit: 1 = TL_MODE_F16, 2 = TL_MODE_F16X2.  src_fp32 bit s: segment s reads raw fp32 rows.
int conv_fwd_ts(const tl_conv_desc& d, cudaStream_t stream, int nsplit, uint32_t src_fp32_mask) {
    static int num_sms[16] = {0};
    int dev = 0;
    TL_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 16) {
        set_error("tl_conv_fwd: device ordinal %d not supported", dev);
        return TL_ERR_UNSUPPORTED;
    }
    if (!num_sms[dev]) {   // per-device one-time setup (function attributes are per device)
        int n = 0;
        TL_CUDA_CHECK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
        TL_CUDA_CHECK(cudaFuncSetAttribute(ts::k_conv_ts<1, ts::NTW1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        TL_CUDA_CHECK(cudaFuncSetAttribute(ts::k_conv_ts<2, ts::NTW2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        num_sms[dev] = n;
    }
    const int sms = num_sms[dev];
    const int n = d.c_out;
    ts::Launch P;
    memset(&P, 0, sizeof(P));
    P.debug = env_int_ts("TL_TS_DEBUG", 0);
    P.sleep_ns = env_int_ts("TL_TS_SLEEP", 64);
    P.num_tiles = (d.n_out + tc::BM - 1) / tc::BM;
    P.chunks_total = 0;
    int idx_rows = 0;
    for (int s = 0; s < d.n_seg; ++s) {
        P.chunks_total += d.seg[s].n_off * (d.seg[s].c_in / 32);
        P.src_fp32[s] = (src_fp32_mask >> s) & 1u;
        P.idx_base[s] = 0, P.idx_owner[s] = 0;
        if (!d.seg[s].index) continue;
        int alias = -1;
        for (int t = 0; t < s; ++t)
            if (d.seg[t].index == d.seg[s].index && d.seg[t].index_stride == d.seg[s].index_stride && d.seg[t].n_off == d.seg[s].n_off)
                alias = t;
        if (alias >= 0) P.idx_base[s] = P.idx_base[alias];
        else P.idx_base[s] = idx_rows, P.idx_owner[s] = 1, idx_rows += d.seg[s].n_off;
    }
    // split-K for layers with few row tiles (deep levels)
    P.splits = 1;
    P.splitk_ws = nullptr;
    const int split_target = env_int_ts("TL_TC_SPLIT_WAVES", 1);
    if (d.splitk_ws && P.num_tiles < split_target * sms) {
        int want = (split_target * sms + P.num_tiles - 1) / P.num_tiles;
        if (want > P.chunks_total / 4) want = P.chunks_total / 4;
        if (want > 1) {
            P.splits = want;
            P.splitk_ws = d.splitk_ws;
            TL_CUDA_CHECK(cudaMemsetAsync(d.splitk_ws, 0, sizeof(float) * (size_t)d.n_out * d.c_out, stream));
        }
    }
    const int num_work = P.num_tiles * P.splits;
    // TMEM: accumulators first, the A ring behind them
    const int chunk_cols = 16 * nsplit;
    P.acc_stride = n;
    P.acc_bufs = (num_work > sms && 2 * n + 8 * chunk_cols <= ts::TMEM_COLS) ? 2 : 1;
    if (env_int_ts("TL_TS_ACC_BUFS", 0)) P.acc_bufs = env_int_ts("TL_TS_ACC_BUFS", 0);
    P.a_col0 = P.acc_bufs * n;
    int ring_chunks = (ts::TMEM_COLS - P.a_col0) / chunk_cols;
    const uint32_t slab = (uint32_t)n * 64u * nsplit;
    // weights: resident when the whole layer fits the budget, else one stage per ring chunk
    const uint32_t res_budget = (uint32_t)env_int_ts("TL_TS_RES_KB", 112) * 1024u, ring_budget = (uint32_t)env_int_ts("TL_TS_RING_KB", 128) * 1024u;
    P.resident = ((uint64_t)P.chunks_total * slab <= res_budget) ? 1 : 0;
    if (!P.resident && (uint64_t)ring_chunks * slab > ring_budget) ring_chunks = (int)(ring_budget / slab);
    int S = ring_chunks >= 15 ? 3 : 2;
    S = env_int_ts("TL_TS_S", S);
    if (S < 2) S = 2;
    if (S > ts::MAX_SLOTS) S = ts::MAX_SLOTS;
    int Q = ring_chunks / S;
    if (env_int_ts("TL_TS_Q", 0) > 0 && env_int_ts("TL_TS_Q", 0) < Q) Q = env_int_ts("TL_TS_Q", 0);
    if (Q > 32) Q = 32;   // the MMA warp looks the fill's slabs up one lane per chunk
    if (Q < 1) {
        set_error("tl_conv_fwd(ts): c_out=%d leaves no room for the TMEM A ring", n);
        return TL_ERR_UNSUPPORTED;
    }
    P.S = S, P.Q = Q;
    P.b_bytes = P.resident ? (uint32_t)P.chunks_total * slab : (uint32_t)(S * Q) * slab;
    const size_t smem = ts::smem_bytes(P.b_bytes);
    if (smem > 227 * 1024) {
        set_error("tl_conv_fwd(ts): %zu bytes of shared memory needed", smem);
        return TL_ERR_UNSUPPORTED;
    }
    int grid = num_work < sms ? num_work : sms;
    if (nsplit == 2) ts::k_conv_ts<2, ts::NTW2><<<grid, 32 * (ts::NTW2 + 7), smem, stream>>>(d, P);
    else ts::k_conv_ts<1, ts::NTW1><<<grid, 32 * (ts::NTW1 + 7), smem, stream>>>(d, P);
    TL_LAUNCH_CHECK();
    if (P.splits > 1) {
        const int64_t total = (int64_t)d.n_out * d.c_out;
        const unsigned eg = (unsigned)((total + 255) / 256 > 1184 ? 1184 : (total + 255) / 256);
        if (nsplit == 2) ts::k_splitk_epilogue<2><<<eg, 256, 0, stream>>>(d, d.splitk_ws);
        else ts::k_splitk_epilogue<1><<<eg, 256, 0, stream>>>(d, d.splitk_ws);
        TL_LAUNCH_CHECK();
    }
    return TL_OK;
}

}  // namespace tl
