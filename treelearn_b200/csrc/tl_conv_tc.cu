// tcgen05 / TMEM segmented gather-GEMM sparse convolution (mode TL_MODE_TF32), sm_100a only.
//
// One CTA owns a tile of 128 output voxels (UMMA M=128, cta_group::1) and all C_out columns
// (UMMA N = C_out, fp32 accumulators in TMEM, double-buffered so the epilogue of tile i overlaps the
// main loop of tile i+1).  The K loop runs over (segment, kernel offset, 32-channel block): per step
//   * 4 producer warps gather the 128 neighbour rows (128 B each, cp.async 16 B through L1, zero-fill for
//     absent neighbours) into a 128B-swizzled K-major A stage and copy the offset's [C_out x 32] weight slab
//     into a B stage -- exactly the canonical SWIZZLE_128B layouts the UMMA shared-memory descriptors name.
//     The ring is deep (S stages, S-2 cp.async groups in flight per thread): the gather is latency bound.
//   * 1 MMA thread issues 4 x tcgen05.mma.kind::tf32 (K=8 each) and tcgen05.commit's the stage back;
//   * 4 epilogue warps tcgen05.ld the accumulator rows, add the residual, and write up to three outputs
//     (raw, and relu(scale*v+shift) for the next layers' BatchNorm+ReLU, rounded to TF32 so the tensor
//     core's operand truncation of those tensors is exact).
// Offsets that no voxel of the tile uses are skipped through the rulebook's per-tile bitmask.
// Layers with few tiles (deep U-Net levels: big weights, few voxels) run split-K: the (offset, k-block) range of a
// tile is divided over several CTAs which red.add their partial accumulators into a zeroed fp32 buffer; a small
// elementwise kernel then applies residual / BN / ReLU.
// Shared memory is kept near 128 KB so that ~96 KB of L1 remains: the ~14x re-read of neighbour rows inside a
// Morton-ordered tile is served by L1, not L2.
// Weights arrive pre-rounded (RN) to TF32, layout [n_off][C_out][C_in] (K-major B operand).
#include <stdlib.h>

#include "tl_common.cuh"

namespace tl {
namespace tc {

constexpr int BM = 128;          // rows per tile == TMEM lanes
constexpr int BK = 32;           // fp32 elements per K block == one 128 B swizzle row
constexpr int MAX_STAGES = 12;
constexpr int A_STAGE_BYTES = BM * 128;
constexpr int kProducerWarps = 8;        // two groups of four; a group fills one whole stage
constexpr int kProducerThreads = 128;    // arrivals per stage (one group)
constexpr int kEpilogueThreads = 128;
constexpr int kMmaWarp = kProducerWarps + 4;
constexpr int kThreads = 32 * (kProducerWarps + 4 + 1);
constexpr int IDX_ROWS = 32;                         // rulebook rows (segment, offset) staged per tile
constexpr int IDX_BUF_BYTES = IDX_ROWS * BM * 4;     // one tile's worth; double buffered

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ int ld_shared_i32(uint32_t addr) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// wait until at most `n` of this thread's cp.async groups are pending (n is warp-uniform)
__device__ __forceinline__ void cp_async_wait_dyn(int n) {
    switch (n) {
        case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
        case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
        case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
        case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
        case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
        case 5: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
        case 6: asm volatile("cp.async.wait_group 6;" ::: "memory"); break;
        case 7: asm volatile("cp.async.wait_group 7;" ::: "memory"); break;
        case 8: asm volatile("cp.async.wait_group 8;" ::: "memory"); break;
        case 9: asm volatile("cp.async.wait_group 9;" ::: "memory"); break;
        default: asm volatile("cp.async.wait_group 10;" ::: "memory"); break;
    }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 (ignored for swizzled K-major, set 1) | [32,46) SBO >> 4 (8 rows
//   x 128 B = 1024) | [46,48) version = 1 (sm100) | [49,52) base offset = 0 (1024 B aligned stages) |
//   [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// cute::UMMA::InstrDescriptor: c_format F32 (1) @4, a/b format TF32 (2) @7/@10, K-major both, N>>3 @17, M>>4 @24
__device__ __forceinline__ uint32_t make_idesc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float round_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}

struct Launch {   // per-launch scalars (kernel parameter)
    int num_tiles;     // row tiles
    int splits;        // CTAs sharing one tile's K range (1 = fused epilogue)
    int chunks_total;  // (segment, offset, k-block) ordinals per tile, masked ones included
    int stages, lag;
    int tmem_cols, buf_cols;
    float* splitk_ws;  // [n_out, c_out] zeroed fp32 accumulation buffer when splits > 1
    int idx_base[TL_MAX_SEG];   // first prefetch row of each indexed segment (segments sharing a table share rows)
    int idx_owner[TL_MAX_SEG];  // 1 = this segment's table rows are fetched (0 = alias of an earlier segment)
};

struct Layout {  // dynamic shared memory carve-up (1024 B aligned base)
    uint32_t a0, b0, b_stage_bytes;
    uint32_t full0, empty0, tfull0, tempty0, tmem_slot, idx0;
    __device__ __forceinline__ uint32_t idx(uint32_t buf, int row, int col) const {
        return idx0 + buf * IDX_BUF_BYTES + (uint32_t)(row * BM + col) * 4u;
    }
    __device__ __forceinline__ uint32_t a(uint32_t s) const { return a0 + s * A_STAGE_BYTES; }
    __device__ __forceinline__ uint32_t b(uint32_t s) const { return b0 + s * b_stage_bytes; }
    __device__ __forceinline__ uint32_t full(uint32_t s) const { return full0 + 8 * s; }
    __device__ __forceinline__ uint32_t empty(uint32_t s) const { return empty0 + 8 * s; }
    __device__ __forceinline__ uint32_t tfull(uint32_t b) const { return tfull0 + 8 * b; }
    __device__ __forceinline__ uint32_t tempty(uint32_t b) const { return tempty0 + 8 * b; }
};

__device__ __forceinline__ Layout carve(uint32_t base, int n, int stages) {
    Layout L;
    L.a0 = base;
    L.b0 = base + stages * A_STAGE_BYTES;
    L.b_stage_bytes = n * 128;
    L.idx0 = L.b0 + stages * L.b_stage_bytes;
    uint32_t off = L.idx0 + 2 * IDX_BUF_BYTES;
    L.full0 = off;
    L.empty0 = off + 8 * MAX_STAGES;
    L.tfull0 = off + 16 * MAX_STAGES;
    L.tempty0 = L.tfull0 + 16;
    L.tmem_slot = L.tfull0 + 32;
    return L;
}
static inline size_t smem_bytes(int n, int stages) {
    return 1024 + (size_t)stages * (A_STAGE_BYTES + (size_t)n * 128) + 2 * IDX_BUF_BYTES + 16 * MAX_STAGES + 64;
}

__device__ __forceinline__ uint32_t seg_mask(const tl_conv_seg& sg, int64_t tile) {
    if (!sg.index) return 1u;
    const uint32_t all = sg.n_off >= 32 ? 0xffffffffu : ((1u << sg.n_off) - 1u);
    return (sg.tile_mask ? sg.tile_mask[tile] : 0xffffffffu) & all;
}

// Enumerate the live chunks of (tile, split): f(seg, k, kb, new_offset) for ordinals in [lo, hi) whose offset is
// in the tile's mask.  Producer and MMA threads run the identical enumeration.
template <typename F>
__device__ __forceinline__ void for_each_chunk(const tl_conv_desc& d, int tile, int lo, int hi, F&& f) {
    int ord = 0;
    for (int s = 0; s < d.n_seg; ++s) {
        const tl_conv_seg& sg = d.seg[s];
        const uint32_t mask = seg_mask(sg, tile);
        const int kblocks = sg.c_in / BK;
        for (int k = 0; k < sg.n_off; ++k, ord += kblocks) {
            if (ord >= hi) return;
            if (ord + kblocks <= lo || !((mask >> k) & 1u)) continue;
            bool fresh = true;
            for (int kb = 0; kb < kblocks; ++kb) {
                const int o = ord + kb;
                if (o < lo || o >= hi) continue;
                f(s, k, kb, fresh);
                fresh = false;
            }
        }
    }
}

__global__ void __launch_bounds__(kThreads, 1) k_conv_tc(const tl_conv_desc d, const Launch P) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int N = d.c_out;
    const Layout L = carve(base, N, P.stages);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (L.tmem_slot - smem_u32(smem_raw)));
    const uint32_t S = (uint32_t)P.stages;
    const int num_work = P.num_tiles * P.splits;
    const int per_split = (P.chunks_total + P.splits - 1) / P.splits;

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < S; ++s) {
            mbar_init(L.full(s), kProducerThreads);
            mbar_init(L.empty(s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(L.tfull(b), 1);
            mbar_init(L.tempty(b), kEpilogueThreads);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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

    if (warp < kProducerWarps) {
        // ===================== producers: gather A rows + copy B slab ==============================
        // Two groups of 4 warps take alternate chunks (the loop is issue/latency bound per warp, so two warps per
        // SM sub-partition double the rate).  Everything loop-invariant is hoisted: per-thread swizzled destination
        // offsets, per-(segment, offset) row pointers; ring position is tracked incrementally (no div/mod).
        const int group = warp >> 2, gw = warp & 3;       // producer group, warp within group
        const int ptid = threadIdx.x & 127;               // thread within group
        const int chunk = lane & 7, sub = lane >> 3;
        const int col = gw * 32 + lane;                   // the tile row whose rulebook entry this thread stages
        uint32_t a_off[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int rl = i * 4 + sub;
            a_off[i] = (uint32_t)((gw * 32 + rl) * 128 + ((chunk ^ (rl & 7)) << 4));
        }
        const int brow = ptid >> 3;                       // B rows brow + 16 j
        const uint32_t b_off0 = (uint32_t)(brow * 128 + ((chunk ^ (brow & 7)) << 4));
        const int nb = N >> 4;
        const uint32_t lag = (uint32_t)P.lag;

        auto prefetch_idx = [&](int tile, uint32_t buf) {   // group 0 only
            for (int s = 0; s < d.n_seg; ++s) {
                const tl_conv_seg& sg = d.seg[s];
                if (!sg.index || !P.idx_owner[s]) continue;
                const int32_t* ip = sg.index + (int64_t)tile * BM + col;
                for (int k = 0; k < sg.n_off; ++k) cp_async4(L.idx(buf, P.idx_base[s] + k, col), ip + (int64_t)k * sg.index_stride);
            }
            cp_async_commit();
        };

        uint32_t slot = 0, phase = 0;          // ring position of the next chunk (all chunks, both groups)
        uint32_t cidx = 0;                     // chunk ordinal in this CTA's stream: owner = cidx & 1
        uint32_t own_issued = 0, own_pub = 0;  // this group's chunks issued / published
        uint32_t pub_slot = (uint32_t)group % S;   // ring slot of this group's next chunk to publish
        uint32_t witer = 0, since_prefetch = 0;
        if (group == 0 && (int)blockIdx.x < num_work) prefetch_idx((int)blockIdx.x / P.splits, 0);
        for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++witer) {
            const int tile = w / P.splits, split = w - tile * P.splits;
            const int lo = split * per_split, hi = min(lo + per_split, P.chunks_total);
            const int64_t row0 = (int64_t)tile * BM;
            const uint32_t ibuf = witer & 1u;
            if (group == 0) cp_async_wait_dyn((int)min(since_prefetch, lag));   // this tile's rulebook rows landed
            since_prefetch = 0;
            asm volatile("bar.sync 1, 256;" ::: "memory");                       // ... and are visible to group 1
            const int wn = w + gridDim.x;
            if (group == 0 && wn < num_work) prefetch_idx(wn / P.splits, ibuf ^ 1u);
            int ord = 0;
            for (int s = 0; s < d.n_seg; ++s) {
                const tl_conv_seg& sg = d.seg[s];
                const uint32_t mask = seg_mask(sg, tile);
                const int kblocks = sg.c_in / BK;
                const float* wseg = sg.weight + (int64_t)brow * sg.c_in + chunk * 4;
                const int64_t wrow_stride = (int64_t)16 * sg.c_in;
                for (int k = 0; k < sg.n_off; ++k, ord += kblocks) {
                    if (ord >= hi) break;
                    if (ord + kblocks <= lo || !((mask >> k) & 1u)) continue;
                    const int kb_lo = max(lo - ord, 0), kb_hi = min(hi - ord, kblocks);
                    // which of this offset's chunks does my group own?
                    const uint32_t first_owned = (cidx & 1u) == (uint32_t)group ? 0u : 1u;
                    if ((int)first_owned < kb_hi - kb_lo) {
                        int my_row;
                        if (sg.index) my_row = ld_shared_i32(L.idx(ibuf, P.idx_base[s] + k, col));
                        else my_row = (row0 + col) < d.n_out ? (int)(row0 + col) : -1;
                        const float* rp[8];
                        uint32_t vmask = 0;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int r = __shfl_sync(0xffffffffu, my_row, i * 4 + sub);
                            vmask |= (r >= 0 ? 1u : 0u) << i;
                            rp[i] = sg.src + (int64_t)max(r, 0) * sg.src_stride + chunk * 4;
                        }
                        const float* wk = wseg + (int64_t)k * N * sg.c_in;
                        for (int kb = kb_lo; kb < kb_hi; ++kb) {
                            const bool mine = (cidx & 1u) == (uint32_t)group;
                            if (mine) {
                                mbar_wait(L.empty(slot), phase ^ 1u);
                                const uint32_t a_st = L.a(slot), b_st = L.b(slot) + b_off0;
#pragma unroll
                                for (int i = 0; i < 8; ++i)
                                    cp_async16(a_st + a_off[i], rp[i] + kb * BK, ((vmask >> i) & 1u) ? 16u : 0u);
                                const float* wp = wk + kb * BK;
                                for (int j = 0; j < nb; ++j) cp_async16(b_st + j * 2048, wp + j * wrow_stride, 16u);
                                cp_async_commit();
                                ++own_issued;
                                ++since_prefetch;
                                if (own_issued - own_pub > lag) {
                                    cp_async_wait_dyn((int)lag);
                                    fence_proxy_async();
                                    mbar_arrive(L.full(pub_slot));
                                    ++own_pub;
                                    pub_slot += 2;
                                    if (pub_slot >= S) pub_slot -= S;
                                }
                            }
                            ++cidx;
                            if (++slot == S) slot = 0, phase ^= 1u;
                        }
                    } else {   // nothing owned here: just advance the stream position
                        const int cnt = kb_hi - kb_lo;
                        for (int c = 0; c < cnt; ++c) {
                            ++cidx;
                            if (++slot == S) slot = 0, phase ^= 1u;
                        }
                    }
                }
            }
        }
        cp_async_wait_dyn(0);
        fence_proxy_async();
        for (; own_pub < own_issued; ++own_pub) {
            mbar_arrive(L.full(pub_slot));
            pub_slot += 2;
            if (pub_slot >= S) pub_slot -= S;
        }
    } else if (warp < kProducerWarps + 4) {
        // ===================== epilogue: TMEM -> registers -> global ================================
        const int ew = warp - kProducerWarps;  // == warp % 4: the TMEM lane quarter this warp may touch
        uint32_t titer = 0;
        for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++titer) {
            const int tile = w / P.splits, split = w - tile * P.splits;
            const uint32_t buf = titer & 1u;
            bool any = false;
            for_each_chunk(d, tile, split * per_split, min((split + 1) * per_split, P.chunks_total),
                           [&](int, int, int, bool) { any = true; });
            mbar_wait(L.tfull(buf), (titer >> 1) & 1u);
            tc_fence_after();
            const int64_t row = (int64_t)tile * BM + ew * 32 + lane;
            const bool live = row < d.n_out;
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + buf * P.buf_cols;
            if (P.splits > 1) {
                if (any) {
                    for (int c0 = 0; c0 < N; c0 += 32) {
                        uint32_t acc[32];
                        tmem_ld32(taddr + c0, acc);
                        if (live) {
                            float* wp = P.splitk_ws + row * N + c0;
#pragma unroll
                            for (int j = 0; j < 32; ++j) atomicAdd(wp + j, __uint_as_float(acc[j]));
                        }
                    }
                }
            } else {
                for (int c0 = 0; c0 < N; c0 += 32) {
                    uint32_t acc[32];
                    if (any) tmem_ld32(taddr + c0, acc);
                    else
#pragma unroll
                        for (int j = 0; j < 32; ++j) acc[j] = 0u;
                    if (live) {
                        const int64_t o = row * N + c0;
                        float v[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
                        if (d.residual) {
                            const float4* rp = reinterpret_cast<const float4*>(d.residual + o);
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float4 r4 = __ldg(rp + j);
                                v[4 * j] += r4.x, v[4 * j + 1] += r4.y, v[4 * j + 2] += r4.z, v[4 * j + 3] += r4.w;
                            }
                        }
                        if (d.out_raw) {
                            float4* op = reinterpret_cast<float4*>(d.out_raw + o);
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                        }
#pragma unroll
                        for (int which = 0; which < 2; ++which) {
                            float* outp = which ? d.out_act2 : d.out_act1;
                            if (!outp) continue;
                            float4* op = reinterpret_cast<float4*>(outp + o);
                            const float4* sp = reinterpret_cast<const float4*>((which ? d.scale2 : d.scale1) + c0);
                            const float4* tp = reinterpret_cast<const float4*>((which ? d.shift2 : d.shift1) + c0);
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float4 s4 = __ldg(sp + j), t4 = __ldg(tp + j);
                                op[j] = make_float4(round_tf32(fmaxf(fmaf(v[4 * j], s4.x, t4.x), 0.f)),
                                                    round_tf32(fmaxf(fmaf(v[4 * j + 1], s4.y, t4.y), 0.f)),
                                                    round_tf32(fmaxf(fmaf(v[4 * j + 2], s4.z, t4.z), 0.f)),
                                                    round_tf32(fmaxf(fmaf(v[4 * j + 3], s4.w, t4.w), 0.f)));
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(L.tempty(buf));
        }
    } else {
        // ===================== MMA issuer (one elected thread) ======================================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(N);
            uint32_t it = 0, titer = 0;
            for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++titer) {
                const int tile = w / P.splits, split = w - tile * P.splits;
                const uint32_t buf = titer & 1u;
                mbar_wait(L.tempty(buf), ((titer >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + buf * P.buf_cols;
                uint32_t first = 1;
                for_each_chunk(d, tile, split * per_split, min((split + 1) * per_split, P.chunks_total),
                               [&](int, int, int, bool) {
                    const uint32_t slot = it % S;
                    mbar_wait(L.full(slot), (it / S) & 1u);
                    tc_fence_after();
                    const uint64_t adesc = make_smem_desc(L.a(slot));
                    const uint64_t bdesc = make_smem_desc(L.b(slot));
#pragma unroll
                    for (int kk = 0; kk < BK / 8; ++kk)  // UMMA_K = 8 for tf32: advance 32 B inside the swizzle atom
                        umma_tf32(tmem_d, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc,
                                  (first && kk == 0) ? 0u : 1u);
                    first = 0;
                    umma_commit(L.empty(slot));
                    ++it;
                });
                if (first) mbar_arrive(L.tfull(buf));  // no pair in this work item: nothing to accumulate
                else umma_commit(L.tfull(buf));
            }
        }
        __syncwarp();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(P.tmem_cols)
                     : "memory");
    }
}

// split-K second pass: v = ws (+ residual) -> raw / act outputs
__global__ void k_splitk_epilogue(const tl_conv_desc d, const float* __restrict__ ws) {
    const int64_t total = (int64_t)d.n_out * d.c_out;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int col = (int)(e % d.c_out);
        float v = ws[e];
        if (d.residual) v += __ldg(d.residual + e);
        if (d.out_raw) d.out_raw[e] = v;
        if (d.out_act1) d.out_act1[e] = round_tf32(fmaxf(fmaf(v, __ldg(d.scale1 + col), __ldg(d.shift1 + col)), 0.f));
        if (d.out_act2) d.out_act2[e] = round_tf32(fmaxf(fmaf(v, __ldg(d.scale2 + col), __ldg(d.shift2 + col)), 0.f));
    }
}

// ------------------------------------------------------------------------------------------------
// 4-channel input convolution (tree_learn.py:37-39): K = 4 is below one UMMA K block, so it runs as SIMT:
// one thread per voxel, 27 gathered float4 rows, weights [27][4][32] broadcast from shared memory.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_conv_in4(const tl_conv_desc d) {
    __shared__ __align__(16) float ws[27 * 4 * 32];
    const tl_conv_seg& sg = d.seg[0];
    for (int e = threadIdx.x; e < sg.n_off * 4 * 32; e += blockDim.x) ws[e] = sg.weight[e];
    __syncthreads();
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t mask = seg_mask(sg, r / TL_TILE_ROWS);
    if (r >= d.n_out) return;
    float acc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = 0.f;
    for (int k = 0; k < sg.n_off; ++k) {
        if (!((mask >> k) & 1u)) continue;
        const int src = __ldg(sg.index + (int64_t)k * sg.index_stride + r);
        if (src < 0) continue;
        const float4 x = __ldg(reinterpret_cast<const float4*>(sg.src + (int64_t)src * sg.src_stride));
        const float4* w4 = reinterpret_cast<const float4*>(ws + k * 128);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 w0 = w4[j], w1 = w4[8 + j], w2 = w4[16 + j], w3 = w4[24 + j];
            acc[4 * j] += x.x * w0.x + x.y * w1.x + x.z * w2.x + x.w * w3.x;
            acc[4 * j + 1] += x.x * w0.y + x.y * w1.y + x.z * w2.y + x.w * w3.y;
            acc[4 * j + 2] += x.x * w0.z + x.y * w1.z + x.z * w2.z + x.w * w3.z;
            acc[4 * j + 3] += x.x * w0.w + x.y * w1.w + x.z * w2.w + x.w * w3.w;
        }
    }
    const int64_t o = r * 32;
    if (d.residual) {
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] += __ldg(d.residual + o + j);
    }
    if (d.out_raw) {
        float4* op = reinterpret_cast<float4*>(d.out_raw + o);
#pragma unroll
        for (int j = 0; j < 8; ++j) op[j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
    }
#pragma unroll
    for (int which = 0; which < 2; ++which) {
        float* outp = which ? d.out_act2 : d.out_act1;
        if (!outp) continue;
        const float* sc = which ? d.scale2 : d.scale1;
        const float* sh = which ? d.shift2 : d.shift1;
        float4* op = reinterpret_cast<float4*>(outp + o);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            op[j] = make_float4(round_tf32(fmaxf(fmaf(acc[4 * j], sc[4 * j], sh[4 * j]), 0.f)),
                                round_tf32(fmaxf(fmaf(acc[4 * j + 1], sc[4 * j + 1], sh[4 * j + 1]), 0.f)),
                                round_tf32(fmaxf(fmaf(acc[4 * j + 2], sc[4 * j + 2], sh[4 * j + 2]), 0.f)),
                                round_tf32(fmaxf(fmaf(acc[4 * j + 3], sc[4 * j + 3], sh[4 * j + 3]), 0.f)));
    }
}

}  // namespace tc

int conv_fwd_simt(const tl_conv_desc& d, cudaStream_t stream);

static bool tc_eligible(const tl_conv_desc& d) {
    if (d.c_out % 32 != 0 || d.c_out > 256) return false;
    for (int s = 0; s < d.n_seg; ++s)
        if (d.seg[s].c_in % tc::BK != 0 || d.seg[s].src_stride % 4 != 0) return false;
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

int conv_fwd_tc(const tl_conv_desc& d, cudaStream_t stream) {
    if (d.n_seg == 1 && d.seg[0].c_in == 4 && d.c_out == 32 && d.seg[0].index && d.seg[0].src_stride == 4) {
        tc::k_conv_in4<<<(unsigned)((d.n_out + 127) / 128), 128, 0, stream>>>(d);
        TL_LAUNCH_CHECK();
        return TL_OK;
    }
    if (!tc_eligible(d)) return conv_fwd_simt(d, stream);
    static int num_sms = 0, smem_budget = 0, split_target = 0;
    if (!num_sms) {
        int dev = 0;
        TL_CUDA_CHECK(cudaGetDevice(&dev));
        TL_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        TL_CUDA_CHECK(cudaFuncSetAttribute(tc::k_conv_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        smem_budget = env_int("TL_TC_SMEM_KB", 128) * 1024;   // keep the rest of the 228 KB as L1 for the gather
        split_target = env_int("TL_TC_SPLIT_WAVES", 2);        // split-K until work items >= waves * SMs
        TL_CUDA_CHECK(cudaFuncSetAttribute(tc::k_conv_tc, cudaFuncAttributePreferredSharedMemoryCarveout,
                                           env_int("TL_TC_CARVEOUT", 58)));
    }
    const int n = d.c_out;
    tc::Launch P;
    P.buf_cols = 32;
    while (P.buf_cols < n) P.buf_cols <<= 1;
    P.tmem_cols = 2 * P.buf_cols;  // <= 512
    const size_t stage = tc::A_STAGE_BYTES + (size_t)n * 128;
    int stages = (int)((smem_budget - 2048 - 2 * tc::IDX_BUF_BYTES) / stage);
    if (stages < 4) stages = 4;
    if (stages > tc::MAX_STAGES) stages = tc::MAX_STAGES;
    while (tc::smem_bytes(n, stages) > 227 * 1024 && stages > 2) --stages;
    P.stages = stages;
    // two producer groups alternate chunks, each keeps `lag` of its own cp.async groups in flight before publishing
    // the oldest; 2*lag < stages keeps the ring deadlock-free (a group can always publish what the MMA waits for)
    P.lag = (stages - 1) / 2;
    if (P.lag < 1) P.lag = 1;
    const size_t smem = tc::smem_bytes(n, stages);
    P.num_tiles = (d.n_out + tc::BM - 1) / tc::BM;
    P.chunks_total = 0;
    for (int s = 0; s < d.n_seg; ++s) P.chunks_total += d.seg[s].n_off * (d.seg[s].c_in / tc::BK);
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
    if (idx_rows > tc::IDX_ROWS) return conv_fwd_simt_fallback_note(d, stream);
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
    tc::k_conv_tc<<<grid, tc::kThreads, smem, stream>>>(d, P);
    TL_LAUNCH_CHECK();
    if (P.splits > 1) {
        const int64_t total = (int64_t)d.n_out * d.c_out;
        tc::k_splitk_epilogue<<<(unsigned)((total + 255) / 256 > 1184 ? 1184 : (total + 255) / 256), 256, 0, stream>>>(
            d, d.splitk_ws);
        TL_LAUNCH_CHECK();
    }
    return TL_OK;
}

}  // namespace tl
