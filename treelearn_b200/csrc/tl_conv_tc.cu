// tcgen05 / TMEM segmented gather-GEMM sparse convolution (modes TL_MODE_TF32 / TL_MODE_F16), sm_100a only.
//
// One CTA owns a tile of 128 output voxels (UMMA M=128, cta_group::1) and all C_out columns
// (UMMA N = C_out, fp32 accumulators in TMEM, double-buffered so the epilogue of tile i overlaps the
// main loop of tile i+1).  The K loop runs over (segment, kernel offset, 32-channel block) "chunks"; per chunk
//   * 4 producer warps gather the 128 neighbour rows (one 64 B / 128 B swizzle row each; cp.async.cg 16 B, absent
//     neighbours read a global zero row so every copy is a uniform 16 B) into a swizzled K-major A stage;
//   * one weight-loader thread lands the chunk's [C_out x 32] weight slab in the B stage with a single TMA bulk copy
//     (weights are packed per slab with the shared-memory swizzle already applied: sparse.pack_weight_tc);
//     both complete on the stage's mbarrier (cp.async noinc arrivals + expect_tx bytes);
//   * 1 MMA thread issues the tcgen05.mma K steps (kind::tf32 K=8 / kind::f16 K=16) and tcgen05.commit's the stage back;
//   * 4 epilogue warps tcgen05.ld the accumulator rows, transpose them through a per-warp shared-memory tile so that
//     all global traffic is full 128 B lines, add the residual and write up to three outputs (raw fp32, and
//     relu(scale*v+shift) for the next layers' BatchNorm+ReLU in the consumers' operand format: TF32-rounded fp32 or fp16).
// Offsets that no voxel of the tile uses are skipped through the rulebook's per-tile bitmask.
// Layers with few tiles (deep U-Net levels: big weights, few voxels) run split-K: the chunk range of a tile is divided
// over several CTAs which red.add their partial accumulators into a zeroed fp32 buffer; a small elementwise kernel
// then applies residual / BN / ReLU.
// Measured limits that shaped this (profiles/r01_ncu_full_k_conv_tc_s2.txt): the LSU data pipe (cp.async shared-memory
// writes + uncoalesced epilogue accesses) was 72 % busy and the producer loop issued 178 instructions per chunk.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "tl_common.cuh"
#include "tl_tc_ptx.cuh"

namespace tl {
namespace tc {

constexpr int BK = 32;           // channels per K block: one swizzle row of 128 B (fp32/TF32 operands) or 64 B (fp16)
constexpr int MAX_STAGES = 12;
constexpr int kProducerWarps = 12;       // gather warps: groups of q warps, group g fills the ring slots with ordinal % G == g
constexpr int kEpilogueThreads = 128;
constexpr int kFirstEpilogueWarp = kProducerWarps;      // 12..15: warp % 4 == TMEM lane quarter
constexpr int kMmaWarp = kProducerWarps + 4;            // 16
constexpr int kSchedWarp = kProducerWarps + 5;          // 17: builds each work item's descriptor (rulebook rows + chunk list)
constexpr int kWeightWarp = kProducerWarps + 6;         // 18: one thread bulk-copies (TMA) each chunk's weight slab
constexpr int kThreads = 32 * (kProducerWarps + 4 + 3);
constexpr int MAX_CHUNKS = 448;                         // live (segment, offset, k-block) entries per work item
constexpr int IDX_ROWS = 28;                         // rulebook rows (segment, offset) staged per tile (3^3 table + spare)
constexpr int IDX_BUF_BYTES = IDX_ROWS * BM * 4 + MAX_CHUNKS * 4 + 64;   // rulebook rows + chunk list + count
constexpr int NDESC = 3;                             // work-item descriptors in flight: the scheduler runs NDESC-1 items ahead
constexpr int EPI_BYTES = 4 * 32 * 128;              // per epilogue warp: 32 rows x 32 fp32 columns, transposed for coalescing
constexpr int SEGTAB_BYTES = 32 * TL_MAX_SEG;

// 256 zero regions of 1 KB: source of absent neighbour rows in "zero row" gather mode (never written).  Spread out so the
// reads do not serialise on one L2 line (a single shared zero row made the kernel 8x slower).
__device__ float g_zero_rows[256 * 256];

// ---- optional timeline trace (TL_TC_DEBUG bit 32): CTA 0 records clock64() at the ring hand-offs of its first fills so that
// the slot round trip can be split into issue / data latency / MMA / barrier hops (tools/trace_conv.py)
constexpr int TRACE_ROLES = 8, TRACE_LEN = 2048;
__device__ unsigned long long g_trace[TRACE_ROLES * TRACE_LEN];
// Compiled in only with -DTL_TC_TRACE=1 (make TRACE=1): the hooks cost registers in the gather loop otherwise.
#ifdef TL_TC_TRACE
constexpr bool kTrace = true;
#else
constexpr bool kTrace = false;
#endif
__device__ __forceinline__ void trace(bool on, int role, uint32_t& pos, uint32_t tag) {
    if (kTrace && on && pos < TRACE_LEN) {
        g_trace[role * TRACE_LEN + pos] = ((unsigned long long)tag << 48) | ((unsigned long long)clock64() & 0xffffffffffffull);
        ++pos;
    }
}


struct Launch {   // per-launch scalars (kernel parameter)
    int num_tiles;     // row tiles
    int splits;        // CTAs sharing one tile's K range (1 = fused epilogue)
    int chunks_total;  // (segment, offset, k-block) ordinals per tile, masked ones included
    int stages;        // ring slots; every slot holds `q` chunks (sub-stages) handed over with ONE barrier round trip
    int q;
    int tmem_cols, buf_cols;   // TMEM columns allocated / per accumulator buffer (two buffers)
    int debug;                 // timing experiments only: 1 = skip MMAs, 2 = skip A gather, 4 = skip epilogue memory ops, 8 = skip B copy
    int acc_ways, acc_cols;    // independent accumulators per buffer (K-step kk -> accumulator kk % ways) and their stride
    float* splitk_ws;  // [n_out, c_out] zeroed fp32 accumulation buffer when splits > 1
    int groups;        // active producer groups (power of two <= kMaxGroups, < stages)
    int use_cg;        // gather A rows with cp.async.cg (bypass L1) instead of .ca
    int sleep_ns;      // back-off of the per-tile barrier waits (0 = spin)
    int prefetch;      // 1: the scheduler bulk-prefetches (L2) the tile's own source rows of submanifold / identity segments
    int zero_row;      // 1: absent neighbours read a zero row (every copy a uniform 16 B); 0: cp.async zero-fill (src-size 0)
    int idx_base[TL_MAX_SEG];   // first prefetch row of each indexed segment (segments sharing a table share rows); -1 = identity
    int idx_owner[TL_MAX_SEG];  // 1 = this segment's table rows are fetched (0 = alias of an earlier segment)
};

struct Layout {  // dynamic shared memory carve-up (1024 B aligned base)
    uint32_t a0, b0, a_stage_bytes, b_stage_bytes;
    uint32_t full0, empty0, tfull0, tempty0, tmem_slot, idx0, epi0, segtab;
    __device__ __forceinline__ uint32_t idx(uint32_t buf, int row, int col) const {
        return idx0 + buf * IDX_BUF_BYTES + (uint32_t)(row * BM + col) * 4u;
    }
    __device__ __forceinline__ uint32_t a(uint32_t sub) const { return a0 + sub * a_stage_bytes; }   // sub = slot * q + i
    __device__ __forceinline__ uint32_t b(uint32_t sub) const { return b0 + sub * b_stage_bytes; }
    __device__ __forceinline__ uint32_t full(uint32_t s) const { return full0 + 8 * s; }
    __device__ __forceinline__ uint32_t empty(uint32_t s) const { return empty0 + 8 * s; }
    __device__ __forceinline__ uint32_t tfull(uint32_t b) const { return tfull0 + 8 * b; }
    __device__ __forceinline__ uint32_t tempty(uint32_t b) const { return tempty0 + 8 * b; }
    __device__ __forceinline__ uint32_t wfull(uint32_t b) const { return tempty0 + 16 + 8 * b; }              // NDESC
    __device__ __forceinline__ uint32_t wempty(uint32_t b) const { return tempty0 + 16 + 8 * NDESC + 8 * b; }  // NDESC
    __device__ __forceinline__ uint32_t list(uint32_t buf, int j) const {
        return idx0 + buf * IDX_BUF_BYTES + IDX_ROWS * BM * 4 + (uint32_t)j * 4u;
    }
    __device__ __forceinline__ uint32_t count(uint32_t buf) const {
        return idx0 + buf * IDX_BUF_BYTES + IDX_ROWS * BM * 4 + MAX_CHUNKS * 4;
    }
};

__device__ __forceinline__ Layout carve(uint32_t base, int n, int substages, int row_bytes) {
    Layout L;
    L.a0 = base;
    L.a_stage_bytes = BM * row_bytes;
    L.b0 = base + substages * L.a_stage_bytes;
    L.b_stage_bytes = n * row_bytes;
    L.idx0 = L.b0 + substages * L.b_stage_bytes;
    L.epi0 = L.idx0 + NDESC * IDX_BUF_BYTES;
    uint32_t off = L.epi0 + EPI_BYTES;
    L.full0 = off;
    L.empty0 = off + 8 * MAX_STAGES;
    L.tfull0 = off + 16 * MAX_STAGES;
    L.tempty0 = L.tfull0 + 16;
    L.tmem_slot = L.tfull0 + 32 + 16 * NDESC;
    L.segtab = L.tfull0 + 64 + 16 * NDESC;
    return L;
}
static inline size_t smem_bytes(int n, int substages, int row_bytes) {
    return 1024 + (size_t)substages * ((size_t)BM * row_bytes + (size_t)n * row_bytes) + NDESC * IDX_BUF_BYTES + EPI_BYTES +
           16 * MAX_STAGES + 64 + 16 * NDESC + SEGTAB_BYTES + 32;
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

// One warp gathers the 128 rows of a chunk: lane (sub, chunk) copies the 16 B piece `chunk` of rows sub, sub + RPI, ...
// Rulebook entries come from the staged table in batches of 8 (independent shared-memory loads first, copies after).
// MODE 0: cp.async.cg with zero-fill for absent rows, 1: cp.async.ca with zero-fill, 2: cp.async.cg, absent rows read zeros.
template <int ROW, int MODE>
__device__ __forceinline__ void gather_rows(uint32_t ia, uint32_t dst0, uint32_t a_off_even, uint32_t a_off_odd, uint64_t src0,
                                            uint32_t seg_stride, uint64_t zsrc0, uint32_t zr0) {
    constexpr int RPI = 32 / (ROW / 16), NR = BM / RPI, BATCH = 8;
#pragma unroll
    for (int i0 = 0; i0 < NR; i0 += BATCH) {
        int r[BATCH];
#pragma unroll
        for (int b = 0; b < BATCH; ++b) r[b] = ld_shared_i32(ia + (uint32_t)((i0 + b) * RPI * 4));
#pragma unroll
        for (int b = 0; b < BATCH; ++b) {
            const int i = i0 + b;
            const uint32_t dst = dst0 + (uint32_t)(i * RPI * ROW) + ((i & 1) ? a_off_odd : a_off_even);
            const uint64_t src = src0 + (uint64_t)(uint32_t)max(r[b], 0) * seg_stride;
            if (MODE == 2)
                cp_async16_cg(dst, reinterpret_cast<const void*>(r[b] >= 0 ? src : zsrc0 + (((zr0 + i) & 255u) << 10)), 16u);
            else if (MODE == 0)
                cp_async16_cg(dst, reinterpret_cast<const void*>(src), r[b] >= 0 ? 16u : 0u);
            else
                cp_async16(dst, reinterpret_cast<const void*>(src), r[b] >= 0 ? 16u : 0u);
        }
    }
}

// Warp roles (608 threads, one CTA per SM, persistent over work items = (row tile, K split)):
//   warps  0..11  producers: G groups of 4 warps; group g gathers the A rows of the chunks whose ordinal in the CTA's
//                 stream is g mod G (4 cp.async of 16 B per thread and chunk; completion via mbarrier noinc arrive)
//   warps 12..15  epilogue (TMEM lane quarter = warp % 4); rows are transposed through a per-warp shared-memory tile so
//                 that every global load/store instruction covers whole 128 B lines
//   warp  16      MMA issuer (lane 0) + TMEM alloc/dealloc
//   warp  17      scheduler: one work item ahead it stages the item's rulebook rows (cp.async) and the list of live
//                 (segment, offset, k-block) chunks in shared memory, so nobody else evaluates masks or split ranges
//   warp  18      weight loader (lane 0): one TMA bulk copy per chunk of the pre-swizzled [C_out x 32] weight slab
// EB: bytes per operand element (4 = fp32 storage / kind::tf32, 2 = fp16 storage / kind::f16); BKC: channels per chunk
// (32, or 64 for fp16 segments whose width is a multiple of 64: 128 B rows = whole cache lines, half the chunks)
template <int EB, int BKC>
__global__ void __launch_bounds__(kThreads, 1) k_conv_tc(const tl_conv_desc d, const Launch P) {
    constexpr int ROW = BKC * EB;         // bytes per operand row in a stage (one swizzle row: 64 or 128 B)
    constexpr int CH = ROW / 16;          // 16 B chunks per row
    constexpr int RPI = 32 / CH;          // rows covered by one warp-wide LDGSTS
    constexpr int KSTEPS = ROW / 32;      // UMMA K steps (32 B each) per stage
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int N = d.c_out;
    const Layout L = carve(base, N, P.stages * P.q, ROW);
    const uint32_t Q = (uint32_t)P.q;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (L.tmem_slot - smem_u32(smem_raw)));
    const uint32_t S = (uint32_t)P.stages;
    const int num_work = P.num_tiles * P.splits;
    const int per_split = (P.chunks_total + P.splits - 1) / P.splits;

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < S; ++s) {
            mbar_init(L.full(s), 32 * P.q + 1);   // the group's gather threads + the weight loader's expect_tx arrive
            mbar_init(L.empty(s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(L.tfull(b), 1);
            mbar_init(L.tempty(b), kEpilogueThreads);
        }
        for (int b = 0; b < NDESC; ++b) {
            mbar_init(L.wfull(b), 33);                               // 32 cp.async completions + lane 0
            mbar_init(L.wempty(b), 32 * P.q * P.groups + kEpilogueThreads + 2);   // every reader of the descriptor
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < d.n_seg) {   // per-segment constants the producers / weight loader look up by segment id
        const tl_conv_seg& sg = d.seg[threadIdx.x];
        const uint32_t t = L.segtab + 32 * threadIdx.x;
        const uint64_t src = (uint64_t)sg.src, wgt = (uint64_t)sg.weight;
        st_shared_v4(t, (uint32_t)src, (uint32_t)(src >> 32), (uint32_t)wgt, (uint32_t)(wgt >> 32));
        st_shared_v4(t + 16, (uint32_t)(sg.src_stride * EB), (uint32_t)(sg.c_in / BKC),
                     sg.index ? (uint32_t)P.idx_base[threadIdx.x] : 0xffffffffu, 0u);
    }
    if (warp == kMmaWarp) {  // TMEM allocation is owned by the MMA warp
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(L.tmem_slot),
                     "r"(P.tmem_cols)
                     : "memory");
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
            if (witer >= NDESC) mbar_wait_sleep(L.wempty(buf), ((witer / NDESC) - 1u) & 1u, (uint32_t)P.sleep_ns);   // readers of item witer-NDESC are done
            const int tile = w / P.splits, split = w - tile * P.splits;
            const int lo = split * per_split, hi = min(lo + per_split, P.chunks_total);
            int ord = 0, pos = 0;
            for (int s = 0; s < d.n_seg; ++s) {
                const tl_conv_seg& sg = d.seg[s];
                const int kblocks = sg.c_in / BKC;
                if (P.prefetch && lane == 0 && (!sg.index || sg.n_off == 27) && sg.src_stride == sg.c_in) {
                    // rows [128 t, 128 t + 128) of the source are this tile's centre taps and most of its 3^3 neighbours
                    // (Morton order): pull them into L2 one work item ahead so the row gathers mostly hit L2
                    const int64_t r0 = (int64_t)tile * BM;
                    const int64_t rows = min((int64_t)BM, (int64_t)d.n_out - r0);
                    const char* pa = reinterpret_cast<const char*>(sg.src) + r0 * sg.src_stride * EB;
                    if (rows > 0)
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pa), "r"((uint32_t)(rows * sg.c_in * EB))
                                     : "memory");
                }
                if (sg.index && P.idx_owner[s]) {   // rulebook rows: lane stages 4 tile rows (16 B) per offset
                    const int32_t* ip = sg.index + (int64_t)tile * BM + 4 * lane;
                    for (int k = 0; k < sg.n_off; ++k)
                        cp_async16(L.idx(buf, P.idx_base[s] + k, 4 * lane), ip + (int64_t)k * sg.index_stride, 16u);
                }
                // lane k owns offset k: live k-blocks inside [lo, hi), exclusive prefix over lanes = list position
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
                    if (mypos + t < MAX_CHUNKS)
                        st_shared_u32(L.list(buf, mypos + t), (uint32_t)((s << 8) | (lane << 3) | (kb_lo + t)));
                pos += __shfl_sync(0xffffffffu, incl, 31);
                ord += sg.n_off * kblocks;
            }
            if (lane == 0) st_shared_u32(L.count(buf), (uint32_t)min(pos, MAX_CHUNKS));
            __syncwarp();
            cp_async_mbar_arrive_noinc(L.wfull(buf));
            if (lane == 0) mbar_arrive(L.wfull(buf));
        }
    } else if (warp < P.q * P.groups) {
        // ===================== producers: gather A rows =============================================
        // A group = q warps fills one ring slot (q chunks): warp gw gathers ALL 128 rows of chunk gw, so the per-chunk
        // bookkeeping (list entry, segment constants, barrier wait / arrive) is paid once per 128 rows.
        // A thread never waits for its copies: `cp.async.mbarrier.arrive.noinc` makes the slot's full barrier count
        // this thread once all its prior cp.async have landed (the CUTLASS sm100 cp.async->UMMA hand-off).
        constexpr int NR = BM / RPI;                      // LDGSTS per lane and chunk (16 fp16 / 32 tf32)
        constexpr int BATCH = 8;                          // rulebook entries fetched from shared memory per batch
        const uint32_t group = (uint32_t)warp / Q, gw = (uint32_t)warp % Q;
        const int chunk = lane % CH, sub = lane / CH;
        // 16 B chunk c of row r lives at chunk c ^ (r & 7) (128 B rows, SWIZZLE_128B) or c ^ ((r >> 1) & 3) (64 B, SWIZZLE_64B);
        // rows advance by RPI per copy: the swizzle term is constant (64 B rows, RPI = 8) or alternates (128 B rows, RPI = 4)
        const uint32_t a_off_even = (uint32_t)(sub * ROW + ((ROW == 128 ? (chunk ^ sub) : (chunk ^ ((sub >> 1) & 3))) << 4));
        const uint32_t a_off_odd = (uint32_t)(sub * ROW + ((ROW == 128 ? (chunk ^ (sub + 4)) : (chunk ^ ((sub >> 1) & 3))) << 4));
        const uint64_t zero_src = (uint64_t)g_zero_rows + (uint32_t)(chunk * 16);
        const uint32_t G = (uint32_t)P.groups;
        uint32_t my_next = group;              // ordinal (in this CTA's stream of slot fills) of my group's next fill
        uint32_t c0 = 0;                       // ordinal of the current work item's first slot fill
        uint32_t slot = group, phase = 0;      // ring position of my_next (G <= S)
        uint32_t witer = 0;
        uint32_t prev_s = 0xffffffffu;
        uint64_t seg_src = 0;
        uint32_t seg_stride = 0, seg_idx = 0xffffffffu;
        const bool tr = kTrace && (P.debug & 32) && blockIdx.x == 0 && gw == 0 && lane == 0 && group < 4;
        uint32_t tpos = 0;
        for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++witer) {
            const int tile = w / P.splits;
            const int64_t row0 = (int64_t)tile * BM;
            const uint32_t buf = witer % NDESC;
            mbar_wait_sleep(L.wfull(buf), (witer / NDESC) & 1u, (uint32_t)P.sleep_ns);
            const uint32_t n = ld_shared_u32(L.count(buf));
            const uint32_t nfill = (n + Q - 1) / Q;
            while (my_next < c0 + nfill) {
                const uint32_t j = (my_next - c0) * Q + gw;     // my chunk of this fill (if j < n)
                const bool have = j < n;
                uint32_t e = 0;
                if (have) e = ld_shared_u32(L.list(buf, (int)j));
                const uint32_t s = e >> 8, k = (e >> 3) & 31u, kb = e & 7u;
                if (have && s != prev_s) {     // segment constants (rarely changes)
                    prev_s = s;
                    const uint2 sa = ld_shared_u2(L.segtab + 32 * s);
                    const uint2 sb = ld_shared_u2(L.segtab + 32 * s + 16);
                    seg_src = ((uint64_t)sa.y << 32 | sa.x) + (uint32_t)(chunk * 16);
                    seg_stride = sb.x;
                    seg_idx = ld_shared_u32(L.segtab + 32 * s + 24);
                }
                trace(tr, (int)group, tpos, (my_next << 2) | 0u);          // begins waiting for the slot
                mbar_wait(L.empty(slot), phase ^ 1u);
                trace(tr, (int)group, tpos, (my_next << 2) | 1u);          // slot is free
                if (have && !(P.debug & 2)) {
                    const uint64_t src0 = seg_src + kb * ROW;
                    const uint32_t dst0 = L.a(slot * Q + gw);
                    // absent neighbours: either a zero-filled copy (src-size 0) or a read of one of 256 spread-out zero rows
                    const uint64_t zsrc0 = zero_src + kb * ROW;
                    const uint32_t zr0 = (uint32_t)tile * 37u + (uint32_t)sub * 16u;
                    if (seg_idx != 0xffffffffu) {
                        const uint32_t ia = L.idx(buf, (int)(seg_idx + k), sub);
                        if (P.zero_row)
                            gather_rows<ROW, 2>(ia, dst0, a_off_even, a_off_odd, src0, seg_stride, zsrc0, zr0);
                        else if (P.use_cg)
                            gather_rows<ROW, 0>(ia, dst0, a_off_even, a_off_odd, src0, seg_stride, zsrc0, zr0);
                        else
                            gather_rows<ROW, 1>(ia, dst0, a_off_even, a_off_odd, src0, seg_stride, zsrc0, zr0);
                    } else {   // identity segment (1x1 projection / residual input): row = tile row
#pragma unroll 4
                        for (int i = 0; i < NR; ++i) {
                            const int64_t row = row0 + i * RPI + sub;
                            const uint32_t dst = dst0 + (uint32_t)(i * RPI * ROW) + ((i & 1) ? a_off_odd : a_off_even);
                            const uint64_t src = src0 + (uint64_t)(uint32_t)(row < d.n_out ? row : 0) * seg_stride;
                            cp_async16_cg(dst, reinterpret_cast<const void*>(src), row < d.n_out ? 16u : 0u);
                        }
                    }
                }
                cp_async_mbar_arrive_noinc(L.full(slot));
                trace(tr, (int)group, tpos, (my_next << 2) | 2u);          // copies issued
                my_next += G;
                slot += G;
                if (slot >= S) slot -= S, phase ^= 1u;
            }
            c0 += nfill;
            mbar_arrive(L.wempty(buf));
        }
    } else if (warp == kWeightWarp) {
        // ===================== weight loader: one TMA bulk copy per chunk ==========================
        // weights are packed [n_off][c_in/32][C_out][32] with the 16 B chunks of every row already swizzled, so the
        // slab of a chunk is one contiguous C_out x ROW block that lands in the B stage as-is
        {
            const uint32_t slab = (uint32_t)N * ROW;
            uint64_t wbase[TL_MAX_SEG];
            uint32_t wkb[TL_MAX_SEG];
#pragma unroll
            for (int s = 0; s < TL_MAX_SEG; ++s) {
                wbase[s] = s < d.n_seg ? (uint64_t)d.seg[s].weight : 0;
                wkb[s] = s < d.n_seg ? (uint32_t)(d.seg[s].c_in / BKC) : 1u;
            }
            uint32_t slot = 0, phase = 0, witer = 0;
            for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++witer) {
                const uint32_t buf = witer % NDESC;
                mbar_wait_sleep(L.wfull(buf), (witer / NDESC) & 1u, (uint32_t)P.sleep_ns);
                const uint32_t n = ld_shared_u32(L.count(buf));
                for (uint32_t j0 = 0; j0 < n; j0 += Q) {
                    const uint32_t cnt = min(Q, n - j0);
                    // lane qi prepares chunk qi's slab address while the slot drains; one elected lane then issues the copies
                    uint64_t wsrc = 0;
                    if ((uint32_t)lane < cnt) {
                        const uint32_t e = ld_shared_u32(L.list(buf, (int)(j0 + lane)));
                        const uint32_t s = e >> 8, k = (e >> 3) & 31u, kb = e & 7u;
                        const uint64_t wb = s == 0 ? wbase[0] : (s == 1 ? wbase[1] : wbase[2]);
                        const uint32_t kbs = s == 0 ? wkb[0] : (s == 1 ? wkb[1] : wkb[2]);
                        wsrc = wb + (uint64_t)(k * kbs + kb) * slab;
                    }
                    mbar_wait(L.empty(slot), phase ^ 1u);
                    if (P.debug & 8) {
                        if (lane == 0) mbar_arrive(L.full(slot));
                    } else {
                        if (lane == 0) mbar_arrive_expect_tx(L.full(slot), cnt * slab);
                        __syncwarp();
                        if ((uint32_t)lane < cnt)
                            bulk_g2s(L.b(slot * Q + lane), reinterpret_cast<const void*>(wsrc), slab, L.full(slot));
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
        const int ew = warp - kFirstEpilogueWarp;  // == warp % 4: the TMEM lane quarter this warp may touch
        const uint32_t epi = L.epi0 + (uint32_t)ew * (32 * 128);
        const int cc = lane & 7, rsub = lane >> 3;  // after the transpose: lane holds 4 columns (cc) of row i*4 + rsub
        uint32_t titer = 0;
        for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++titer) {
            const int tile = w / P.splits;
            const uint32_t buf = titer & 1u;          // TMEM accumulator buffer
            const uint32_t dbuf = titer % NDESC;      // work-item descriptor buffer
            mbar_wait_sleep(L.wfull(dbuf), (titer / NDESC) & 1u, (uint32_t)P.sleep_ns);
            const bool any = ld_shared_u32(L.count(dbuf)) != 0u;
            mbar_arrive(L.wempty(dbuf));
            const int64_t wrow0 = (int64_t)tile * BM + ew * 32;   // first row of this warp's quarter
            // residual rows of the first 32-column block are fetched BEFORE waiting for the accumulator: their DRAM
            // latency (8 dependent ~1 us loads per tile when issued inside the store loop) hides behind the main loop
            const bool has_res = d.residual != nullptr && P.splits == 1 && !(P.debug & 4);
            float4 res[8];
            auto fetch_residual = [&](int c0) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int64_t grow = wrow0 + i * 4 + rsub;
                    res[i] = (has_res && grow < d.n_out)
                                 ? __ldg(reinterpret_cast<const float4*>(d.residual + grow * N + c0 + 4 * cc))
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            };
            fetch_residual(0);
            mbar_wait_sleep(L.tfull(buf), (titer >> 1) & 1u, (uint32_t)P.sleep_ns);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + buf * P.buf_cols;
            if (any || P.splits == 1) {
                for (int c0 = 0; c0 < N; c0 += 32) {
                    if (c0 > 0) fetch_residual(c0);     // in flight during the TMEM load + transpose of this block
                    uint32_t acc[32];
                    if (any) {
                        tmem_ld32(taddr + c0, acc);
                        for (int way = 1; way < P.acc_ways; ++way) {
                            uint32_t more[32];
                            tmem_ld32(taddr + way * P.acc_cols + c0, more);
#pragma unroll
                            for (int j = 0; j < 32; ++j) acc[j] = __float_as_uint(__uint_as_float(acc[j]) + __uint_as_float(more[j]));
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) acc[j] = 0u;
                    }
                    // row `lane` -> shared (16 B chunk j at j ^ (lane & 7): conflict-free both ways)
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        st_shared_v4(epi + lane * 128 + ((j ^ (lane & 7)) << 4), acc[4 * j], acc[4 * j + 1], acc[4 * j + 2],
                                     acc[4 * j + 3]);
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
                        if (P.splits > 1) {   // 16 B vector reductions into the split-K buffer
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(P.splitk_ws + o), "f"(v.x),
                                         "f"(v.y), "f"(v.z), "f"(v.w)
                                         : "memory");
                            continue;
                        }
                        v.x += res[i].x, v.y += res[i].y, v.z += res[i].z, v.w += res[i].w;
                        if (d.out_raw) *reinterpret_cast<float4*>(d.out_raw + o) = v;
                        if (d.out_act1) {
                            const float a0 = fmaxf(fmaf(v.x, s1.x, t1.x), 0.f), a1 = fmaxf(fmaf(v.y, s1.y, t1.y), 0.f);
                            const float a2 = fmaxf(fmaf(v.z, s1.z, t1.z), 0.f), a3 = fmaxf(fmaf(v.w, s1.w, t1.w), 0.f);
                            store_act4<EB>(d.out_act1, grow, N, col, a0, a1, a2, a3);
                        }
                        if (d.out_act2) {
                            const float a0 = fmaxf(fmaf(v.x, s2.x, t2.x), 0.f), a1 = fmaxf(fmaf(v.y, s2.y, t2.y), 0.f);
                            const float a2 = fmaxf(fmaf(v.z, s2.z, t2.z), 0.f), a3 = fmaxf(fmaf(v.w, s2.w, t2.w), 0.f);
                            store_act4<EB>(d.out_act2, grow, N, col, a0, a1, a2, a3);
                        }
                    }
                    __syncwarp();
                }
            }
            tc_fence_before();
            mbar_arrive(L.tempty(buf));
        }
    } else if (warp == kMmaWarp) {
        // ===================== MMA issuer (one elected thread) ======================================
        {
            const uint32_t idesc = make_idesc(N, EB == 2);
            const uint64_t adesc0 = make_smem_desc(L.a0, ROW), bdesc0 = make_smem_desc(L.b0, ROW);
            const uint32_t b_step = L.b_stage_bytes >> 4;
            const uint32_t wmask = (uint32_t)(P.acc_ways - 1);
            uint32_t slot = 0, phase = 0, titer = 0;
            const bool trm = kTrace && (P.debug & 32) && blockIdx.x == 0 && lane == 0;
            uint32_t tposm = 0, fill_no = 0;
            for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++titer) {
                const uint32_t buf = titer & 1u;          // TMEM accumulator buffer
                const uint32_t dbuf = titer % NDESC;      // work-item descriptor buffer
                mbar_wait_sleep(L.wfull(dbuf), (titer / NDESC) & 1u, (uint32_t)P.sleep_ns);
                const uint32_t n = ld_shared_u32(L.count(dbuf));
                __syncwarp();
                if (elect_one()) mbar_arrive(L.wempty(dbuf));
                mbar_wait_sleep(L.tempty(buf), ((titer >> 1) & 1u) ^ 1u, (uint32_t)P.sleep_ns);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + buf * P.buf_cols;
                for (uint32_t j0 = 0; j0 < n; j0 += Q) {
                    const uint32_t cnt = min(Q, n - j0);
                    trace(trm, 4, tposm, (fill_no << 2) | 0u);             // begins waiting for the data
                    mbar_wait(L.full(slot), phase);
                    trace(trm, 4, tposm, (fill_no << 2) | 1u);             // data landed
                    tc_fence_after();
                    if (elect_one()) {
                        if (!(P.debug & 1)) {
                            constexpr uint32_t A_STEP = (uint32_t)(BM * ROW) >> 4;   // descriptor address units (16 B) per chunk
                            uint64_t adesc = adesc0 + (uint64_t)(slot * Q * A_STEP);
                            uint64_t bdesc = bdesc0 + (uint64_t)(slot * Q * b_step);
                            // chunk 0 of the fill may open the accumulator (first chunk of the work item) ...
#pragma unroll
                            for (int kk = 0; kk < KSTEPS; ++kk) {  // UMMA K step = 32 B (8 tf32 / 16 fp16): advance inside the swizzle atom
                                // K-step kk accumulates into its own TMEM tile kk % ways (the epilogue adds the tiles up)
                                const uint32_t way = (uint32_t)kk & wmask;
                                const uint32_t accum = (j0 == 0 && (uint32_t)kk <= wmask) ? 0u : 1u;
                                if (EB == 4)
                                    umma_tf32(tmem_d + way * P.acc_cols, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, accum);
                                else
                                    umma_f16(tmem_d + way * P.acc_cols, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, accum);
                            }
                            // ... every later chunk accumulates
                            for (uint32_t qi = 1; qi < cnt; ++qi) {
                                adesc += A_STEP;
                                bdesc += b_step;
#pragma unroll
                                for (int kk = 0; kk < KSTEPS; ++kk) {
                                    const uint32_t way = (uint32_t)kk & wmask;
                                    if (EB == 4)
                                        umma_tf32_acc(tmem_d + way * P.acc_cols, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc);
                                    else
                                        umma_f16_acc(tmem_d + way * P.acc_cols, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc);
                                }
                            }
                        }
                        trace(trm, 4, tposm, (fill_no << 2) | 3u);             // MMAs issued
                        umma_commit(L.empty(slot));
                    }
                    __syncwarp();
                    trace(trm, 4, tposm, (fill_no << 2) | 2u);             // MMAs + commit issued
                    ++fill_no;
                    if (++slot == S) slot = 0, phase ^= 1u;
                }
                if (elect_one()) {
                    if (n == 0) mbar_arrive(L.tfull(buf));  // no pair in this work item: nothing to accumulate
                    else umma_commit(L.tfull(buf));
                }
                __syncwarp();
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(P.tmem_cols)
                     : "memory");
    }
}

// split-K second pass: v = ws (+ residual) -> raw / act outputs
template <int EB>
__global__ void k_splitk_epilogue(const tl_conv_desc d, const float* __restrict__ ws) {
    const int64_t total = (int64_t)d.n_out * d.c_out;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int col = (int)(e % d.c_out);
        float v = ws[e];
        if (d.residual) v += __ldg(d.residual + e);
        if (d.out_raw) d.out_raw[e] = v;
        if (d.out_act1) store_act1<EB>(d.out_act1, e, d.c_out, fmaxf(fmaf(v, __ldg(d.scale1 + col), __ldg(d.shift1 + col)), 0.f));
        if (d.out_act2) store_act1<EB>(d.out_act2, e, d.c_out, fmaxf(fmaf(v, __ldg(d.scale2 + col), __ldg(d.shift2 + col)), 0.f));
    }
}

// ------------------------------------------------------------------------------------------------
// 4-channel input convolution (tree_learn.py:37-39): K = 4 is below one UMMA K block, so it runs as SIMT:
// one thread per voxel, 27 gathered float4 rows, weights [27][4][32] broadcast from shared memory.
// ------------------------------------------------------------------------------------------------
// PERM: outputs in the P-layout of the tensor-memory-A kernel (position m of the 32-channel row = logical channel p_chan(m))
template <int EB, bool PERM>
__global__ void __launch_bounds__(128) k_conv_in4(const tl_conv_desc d) {
    __shared__ __align__(16) float ws[27 * 4 * 32];
    __shared__ __align__(16) float stage[4][32 * 32];     // per warp: 32 rows x 32 columns, transposed for coalesced stores
    const tl_conv_seg& sg = d.seg[0];
    for (int e = threadIdx.x; e < sg.n_off * 4 * 32; e += blockDim.x) ws[e] = sg.weight[e];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t wrow0 = (int64_t)blockIdx.x * blockDim.x + warp * 32;
    const int64_t r = wrow0 + lane;
    const bool live = r < d.n_out;
    const uint32_t mask = live ? seg_mask(sg, r / TL_TILE_ROWS) : 0u;
    float acc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = 0.f;
    // offsets in groups of 9: all rulebook entries, then all gathered rows (independent loads in flight), then the FMAs
    for (int g = 0; g < 27; g += 9) {
        int src[9];
        float4 x[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            const int k = g + j;
            src[j] = (k < sg.n_off && ((mask >> k) & 1u)) ? __ldg(sg.index + (int64_t)k * sg.index_stride + r) : -1;
        }
#pragma unroll
        for (int j = 0; j < 9; ++j)
            x[j] = src[j] >= 0 ? __ldg(reinterpret_cast<const float4*>(sg.src + (int64_t)src[j] * sg.src_stride))
                               : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            if (src[j] < 0) continue;
            const float4* w4 = reinterpret_cast<const float4*>(ws + (g + j) * 128);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 w0 = w4[c], w1 = w4[8 + c], w2 = w4[16 + c], w3 = w4[24 + c];
                acc[4 * c] += x[j].x * w0.x + x[j].y * w1.x + x[j].z * w2.x + x[j].w * w3.x;
                acc[4 * c + 1] += x[j].x * w0.y + x[j].y * w1.y + x[j].z * w2.y + x[j].w * w3.y;
                acc[4 * c + 2] += x[j].x * w0.z + x[j].y * w1.z + x[j].z * w2.z + x[j].w * w3.z;
                acc[4 * c + 3] += x[j].x * w0.w + x[j].y * w1.w + x[j].z * w2.w + x[j].w * w3.w;
            }
        }
    }
    // row `lane` -> shared (16 B chunk j at j ^ (lane & 7)), then every store instruction covers whole 128 B lines
    float* st = stage[warp];
#pragma unroll
    for (int j = 0; j < 8; ++j)     // positions 4j .. 4j+3 of the row
        *reinterpret_cast<float4*>(st + lane * 32 + ((j ^ (lane & 7)) << 2)) =
            PERM ? make_float4(acc[p_chan(4 * j)], acc[p_chan(4 * j + 1)], acc[p_chan(4 * j + 2)], acc[p_chan(4 * j + 3)])
                 : make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
    __syncwarp();
    const int cc = lane & 7, rsub = lane >> 3, col = 4 * cc;
    float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), t1 = s1, s2 = s1, t2 = s1;
    auto vec4 = [&](const float* p) {   // scale / shift (logical order) of the 4 channels stored at positions col .. col+3
        return PERM ? make_float4(__ldg(p + p_chan(col)), __ldg(p + p_chan(col + 1)), __ldg(p + p_chan(col + 2)), __ldg(p + p_chan(col + 3)))
                    : __ldg(reinterpret_cast<const float4*>(p + col));
    };
    if (d.out_act1) s1 = vec4(d.scale1), t1 = vec4(d.shift1);
    if (d.out_act2) s2 = vec4(d.scale2), t2 = vec4(d.shift2);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int rr = i * 4 + rsub;
        float4 v = *reinterpret_cast<const float4*>(st + rr * 32 + ((cc ^ (rr & 7)) << 2));
        const int64_t grow = wrow0 + rr;
        if (grow >= d.n_out) continue;
        const int64_t o = grow * 32 + col;
        if (d.residual) {
            const float4 r4 = __ldg(reinterpret_cast<const float4*>(d.residual + o));
            v.x += r4.x, v.y += r4.y, v.z += r4.z, v.w += r4.w;
        }
        if (d.out_raw) *reinterpret_cast<float4*>(d.out_raw + o) = v;
        if (d.out_act1)
            store_act4<EB>(d.out_act1, grow, 32, col, fmaxf(fmaf(v.x, s1.x, t1.x), 0.f), fmaxf(fmaf(v.y, s1.y, t1.y), 0.f),
                           fmaxf(fmaf(v.z, s1.z, t1.z), 0.f), fmaxf(fmaf(v.w, s1.w, t1.w), 0.f));
        if (d.out_act2)
            store_act4<EB>(d.out_act2, grow, 32, col, fmaxf(fmaf(v.x, s2.x, t2.x), 0.f), fmaxf(fmaf(v.y, s2.y, t2.y), 0.f),
                           fmaxf(fmaf(v.z, s2.z, t2.z), 0.f), fmaxf(fmaf(v.w, s2.w, t2.w), 0.f));
    }
}

}  // namespace tc

int conv_fwd_simt(const tl_conv_desc& d, cudaStream_t stream);

static bool tc_eligible(const tl_conv_desc& d) {
    if (d.c_out % 32 != 0 || d.c_out > 256) return false;
    for (int s = 0; s < d.n_seg; ++s)
        if (d.seg[s].c_in % tc::BK != 0 || d.seg[s].c_in > 8 * tc::BK || d.seg[s].src_stride % 4 != 0) return false;
    return true;
}

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

static int conv_fwd_simt_fallback_note(const tl_conv_desc&, cudaStream_t) {
    set_error("tl_conv_fwd(tf32): more than %d distinct rulebook rows per tile are not staged by the tcgen05 path",
              tc::IDX_ROWS);
    return TL_ERR_UNSUPPORTED;
}

// the 4-channel network input is always fp32; the outputs follow the operand format `fmt` (tc::FMT_*) and, with `perm`,
// the P-layout of the tensor-memory-A kernel (raw fp32 output included)
int conv_fwd_in4(const tl_conv_desc& d, cudaStream_t stream, int fmt, bool perm) {
    const unsigned grid = (unsigned)((d.n_out + 127) / 128);
    if (fmt == tc::FMT_F16 && perm) tc::k_conv_in4<tc::FMT_F16, true><<<grid, 128, 0, stream>>>(d);
    else if (fmt == tc::FMT_F16) tc::k_conv_in4<tc::FMT_F16, false><<<grid, 128, 0, stream>>>(d);
    else if (fmt == tc::FMT_F16X2) tc::k_conv_in4<tc::FMT_F16X2, true><<<grid, 128, 0, stream>>>(d);
    else tc::k_conv_in4<tc::FMT_TF32, false><<<grid, 128, 0, stream>>>(d);
    TL_LAUNCH_CHECK();
    return TL_OK;
}

int conv_fwd_tc(const tl_conv_desc& d, cudaStream_t stream, bool half) {
    if (d.n_seg == 1 && d.seg[0].c_in == 4 && d.c_out == 32 && d.seg[0].index && d.seg[0].src_stride == 4)
        return conv_fwd_in4(d, stream, half ? tc::FMT_F16 : tc::FMT_TF32, false);
    if (!tc_eligible(d)) {
        if (half) {
            set_error("tl_conv_fwd(f16): shape not eligible for the tcgen05 path (c_in %% 32, c_out %% 32, c_out <= 256)");
            return TL_ERR_UNSUPPORTED;
        }
        return conv_fwd_simt(d, stream);
    }
    // fp16 convolutions whose every segment is a multiple of 64 channels wide run 64-channel chunks (128 B rows, SWIZZLE_128B);
    // the weights must be packed with the same rule (sparse.pack_weight_tc)
    int bkc = 32;
    if (half && env_int("TL_TC_BK64", 1)) {
        bkc = 64;
        for (int s = 0; s < d.n_seg; ++s)
            if (d.seg[s].c_in % 64 != 0) bkc = 32;
    }
    const int row_bytes = half ? 2 * bkc : 128;
    // per device: SM count and the kernels' dynamic shared-memory attribute (function attributes are per device)
    static int num_sms_dev[16] = {0}, smem_budget = 0, split_target = 0;
    int dev = 0;
    TL_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 16) return TL_ERR_UNSUPPORTED;
    int& num_sms = num_sms_dev[dev];
    if (!num_sms) {
        TL_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        TL_CUDA_CHECK(cudaFuncSetAttribute(tc::k_conv_tc<4, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        TL_CUDA_CHECK(cudaFuncSetAttribute(tc::k_conv_tc<2, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        TL_CUDA_CHECK(cudaFuncSetAttribute(tc::k_conv_tc<2, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        smem_budget = env_int("TL_TC_SMEM_KB", 200) * 1024;   // keep the rest of the 228 KB as L1 for the gather
        split_target = env_int("TL_TC_SPLIT_WAVES", 1);        // split-K until work items >= waves * SMs
    }
    const int n = d.c_out;
    tc::Launch P;
    P.acc_cols = 32;
    while (P.acc_cols < n) P.acc_cols <<= 1;
    P.debug = env_int("TL_TC_DEBUG", 0);
    P.acc_ways = env_int("TL_TC_WAYS", 1);
    while (P.acc_ways > 1 && (2 * P.acc_ways * P.acc_cols > 512 || P.acc_ways > row_bytes / 32)) P.acc_ways >>= 1;
    P.buf_cols = P.acc_ways * P.acc_cols;
    P.tmem_cols = 2 * P.buf_cols;  // <= 512
    // ring = `stages` slots of `q` chunks each (a chunk = 128 gathered rows + the [C_out x 32] weight slab); one barrier
    // round trip hands over a whole slot, so the single-warp MMA / weight-loader loops pay their fixed cost once per q chunks
    const size_t sub = (size_t)(tc::BM + n) * row_bytes;
    const int ring_budget = smem_budget - 2048 - tc::NDESC * tc::IDX_BUF_BYTES - tc::EPI_BYTES;
    // q = chunks per slot (a divisor of the 12 gather warps): the largest of 4, 3, 2, 1 that still leaves `min_slots` slots.
    // Measured (profiles/r01_conv_tc_history.md): bigger slots beat more chunks in flight, and a ring that squeezes the L1
    // below ~25 KB slows the gather down (cp.async misses are tracked in L1), hence the 200 KB default budget.
    int q = env_int("TL_TC_Q", 0);
    const int min_slots = env_int("TL_TC_MIN_SLOTS", 3);
    if (q <= 0 || tc::kProducerWarps % q != 0) {
        q = 1;
        const int cand[4] = {4, 3, 2, 1};
        for (int c = 0; c < 4; ++c)
            if ((int)((size_t)ring_budget / (cand[c] * sub)) >= min_slots) {
                q = cand[c];
                break;
            }
    }
    while (q > 1 && (size_t)ring_budget / (q * sub) < 2) --q;
    while (tc::kProducerWarps % q != 0) --q;
    int stages = (int)(ring_budget / (q * sub));
    if (stages < 2) stages = 2;
    if (stages > tc::MAX_STAGES) stages = tc::MAX_STAGES;
    while (tc::smem_bytes(n, stages * q, row_bytes) > 227 * 1024 && stages > 2) --stages;
    P.stages = stages;
    P.q = q;
    // the producer groups take the CTA's slot fills round-robin; G <= stages keeps the ring deadlock-free
    int groups = env_int("TL_TC_GROUPS", tc::kProducerWarps / q);   // a group = q warps (one per chunk of a slot)
    if (groups > tc::kProducerWarps / q) groups = tc::kProducerWarps / q;
    if (groups > stages) groups = stages;
    if (groups < 1) groups = 1;
    P.groups = groups;
    P.use_cg = env_int("TL_TC_CG", 1);
    P.zero_row = env_int("TL_TC_ZERO_ROW", 1);
    P.prefetch = env_int("TL_TC_PREFETCH", 1);
    P.sleep_ns = env_int("TL_TC_SLEEP", 64);
    const size_t smem = tc::smem_bytes(n, stages * q, row_bytes);
    P.num_tiles = (d.n_out + tc::BM - 1) / tc::BM;
    P.chunks_total = 0;
    for (int s = 0; s < d.n_seg; ++s) P.chunks_total += d.seg[s].n_off * (d.seg[s].c_in / bkc);
    int idx_rows = 0;
    for (int s = 0; s < d.n_seg; ++s) {
        P.idx_base[s] = 0, P.idx_owner[s] = 0;
        if (!d.seg[s].index) continue;
        int alias = -1;
        for (int t = 0; t < s; ++t)
            if (d.seg[t].index == d.seg[s].index && d.seg[t].index_stride == d.seg[s].index_stride &&
                d.seg[t].n_off == d.seg[s].n_off)
                alias = t;
        if (alias >= 0) P.idx_base[s] = P.idx_base[alias];
        else P.idx_base[s] = idx_rows, P.idx_owner[s] = 1, idx_rows += d.seg[s].n_off;
    }
    if (idx_rows > tc::IDX_ROWS || P.chunks_total > tc::MAX_CHUNKS) return conv_fwd_simt_fallback_note(d, stream);
    for (int s = 0; s < d.n_seg; ++s) {
        TL_REQUIRE(!d.seg[s].index || d.seg[s].index_stride % 4 == 0, "tl_conv_fwd(tf32): index_stride must be a multiple of 4");
        TL_REQUIRE(d.seg[s].c_in <= 8 * tc::BK, "tl_conv_fwd(tcgen05): segment %d has c_in=%d > 256 (split it into two segments)",
                   s, d.seg[s].c_in);
    }
    P.splits = 1;
    P.splitk_ws = nullptr;
    if (d.splitk_ws && P.num_tiles < split_target * num_sms) {
        int want = (split_target * num_sms + P.num_tiles - 1) / P.num_tiles;
        if (want > P.chunks_total / 4) want = P.chunks_total / 4;   // keep >= 4 chunks per work item
        if (want > 1) {
            P.splits = want;
            P.splitk_ws = d.splitk_ws;
            TL_CUDA_CHECK(cudaMemsetAsync(d.splitk_ws, 0, sizeof(float) * (size_t)d.n_out * d.c_out, stream));
        }
    }
    int grid = P.num_tiles * P.splits;
    if (grid > num_sms) grid = num_sms;   // 1 CTA per SM (launch bounds); persistent over work items
    if (half && bkc == 64) tc::k_conv_tc<2, 64><<<grid, tc::kThreads, smem, stream>>>(d, P);
    else if (half) tc::k_conv_tc<2, 32><<<grid, tc::kThreads, smem, stream>>>(d, P);
    else tc::k_conv_tc<4, 32><<<grid, tc::kThreads, smem, stream>>>(d, P);
    TL_LAUNCH_CHECK();
    if (P.splits > 1) {
        const int64_t total = (int64_t)d.n_out * d.c_out;
        const unsigned eg = (unsigned)((total + 255) / 256 > 1184 ? 1184 : (total + 255) / 256);
        if (half) tc::k_splitk_epilogue<2><<<eg, 256, 0, stream>>>(d, d.splitk_ws);
        else tc::k_splitk_epilogue<4><<<eg, 256, 0, stream>>>(d, d.splitk_ws);
        TL_LAUNCH_CHECK();
    }
    return TL_OK;
}

}  // namespace tl

// debug: copy the timeline trace of the last TL_TC_DEBUG=32 launch to the host (roles x TRACE_LEN u64 = tag << 48 | clock)
extern "C" int tl_debug_copy_trace(void* host, size_t bytes) {
    const size_t want = sizeof(unsigned long long) * tl::tc::TRACE_ROLES * tl::tc::TRACE_LEN;
    if (bytes < want) return TL_ERR_ARG;
    cudaDeviceSynchronize();
    return cudaMemcpyFromSymbol(host, tl::tc::g_trace, want) == cudaSuccess ? TL_OK : TL_ERR_CUDA;
}
