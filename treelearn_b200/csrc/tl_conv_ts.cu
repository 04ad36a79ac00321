// tcgen05 sparse convolution with the A operand in TENSOR MEMORY (TS-form tcgen05.mma), sm_100a only.
// Modes TL_MODE_F16 (fp16 operands) and TL_MODE_F16X2 (two-term fp16 split of both operands: hi + lo, three MMAs per
// K step = fp32-equivalent products), fp32 accumulation in both.
//
// Measurements that shaped it (profiles/r02_ts_probe.txt, r02_trace_ts_v1_*.txt, r01_conv_tc_history.md):
//   * an M128 x N x K16 fp16 MMA with A in shared memory costs 41 / 49 cycles at N = 32 / 64, with A in tensor memory
//     16.6 / 32.8 (the N/2 floor): the C_out <= 64 layers (80 % of the U-Net's bytes) want the TS form;
//   * every mbarrier hand-off between warp roles costs 100 - 900 cycles; a first TS kernel that kept round 1's roles
//     (16 transfer warps -> 1 MMA warp -> 4 epilogue warps, descriptors from a scheduler warp) spent 12 500 cycles per
//     128-row tile in those hops and was slower than the shared-memory kernel.
// So the CTA (one per SM, persistent) is split into G independent GROUPS of 4 warps.  A group owns one 128-row tile at a
// time, a private fp32 accumulator and a private ring of A chunks in tensor memory, and does everything for its tile:
//   * gather: a quad of lanes reads one contiguous 16 B x 4 piece of a neighbour row straight from global / L1 into
//     registers (8 rows per LDG.128; absent neighbours are predicated off: no traffic), 2-3 chunks ahead of their use
//     (a chunk = 128 rows x 32 channels of one (segment, offset, k-block));
//   * tcgen05.st.16x256b writes the chunk into the group's TMEM ring -- that shape's fragment layout IS "4 lanes per row,
//     8 B each" -- then ONE group barrier (bar.sync, 128 threads) and the group's leader lane issues
//     tcgen05.mma [d_tmem], [a_tmem], b_desc and tcgen05.commit's the ring slot back (the only mbarrier on the path);
//   * epilogue: the group's four warps are the four TMEM lane quarters; tcgen05.ld.16x256b hands every lane 8 channels of
//     4 rows, which go out as 16 B / 32 B vectors that a quad of lanes makes a contiguous row piece: no shared memory.
// Weights: the whole layer stays resident in shared memory when it fits (C = 32: 54 KB), else one weight stream per CTA
// (TMA bulk copies into a ring) feeds all G groups, which then walk the kernel offsets in the same order: weight traffic
// from L2 is paid once per G tiles.
//
// Channel order ("P-layout", treelearn_b200/sparse.py): within every 32-channel block the tensors this kernel reads and
// writes (fp32 residual stream and activated operands) store logical channel 8g + 2q + e at position 8q + 2g + e
// (q = 0..3, g = 0..3, e = 0..1), which is what makes both the gathered 16 B piece of a lane and the 8 accumulator columns
// a lane receives from tcgen05.ld.16x256b contiguous in memory.  The weights carry the matching K order
// (sparse.pack_weight_ts); scale / shift vectors stay in logical order.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "tl_common.cuh"
#include "tl_tc_ptx.cuh"

namespace tl {
namespace ts {

using namespace tl::tc;

constexpr int MAX_R = 16;                  // A chunks per group ring
constexpr int MAX_NB = 64;                 // weight ring slots
constexpr int IDX_ROWS = 28;               // rulebook rows staged per tile and group
constexpr int IDX_BUF_BYTES = IDX_ROWS * BM * 4;
constexpr int TMEM_COLS = 512;
constexpr int MAX_LIST_BYTES = 448 * 4 + 16;   // a tile's chunk list (k_conv_ts::LIST_BYTES)

// ---- optional timeline trace (make TRACE=1, TL_TS_DEBUG bit 32): group 0 of CTA 0 records clock64() (tools/trace_ts.py)
constexpr int TRACE_ROLES = 8, TRACE_LEN = 4096;
__device__ unsigned long long g_trace_ts[TRACE_ROLES * TRACE_LEN];
#ifdef TL_TC_TRACE
constexpr bool kTrace = true;
#else
constexpr bool kTrace = false;
#endif
__device__ __forceinline__ void trace(bool on, int role, uint32_t& pos, uint32_t tag) {
    if (kTrace && on && pos < TRACE_LEN) {
        g_trace_ts[role * TRACE_LEN + pos] = ((unsigned long long)tag << 48) | ((unsigned long long)clock64() & 0xffffffffffffull);
        ++pos;
    }
}

struct Launch {
    int num_tiles, rounds, chunks_total;
    int R;             // A chunks in a group's TMEM ring
    int group_cols;    // TMEM columns per group: [0, c_out) accumulator, then R chunks
    int resident;      // 1: every weight slab of the layer stays in shared memory; 0: weight ring of `nb` slabs
    int nb;
    uint32_t b_bytes;  // shared-memory bytes of the weight region
    int debug;         // timing experiments: 1 skip MMAs, 2 skip row loads, 4 skip epilogue memory ops, 8 skip weight copies, 16 skip tcgen05.st
    int idx_base[TL_MAX_SEG], idx_owner[TL_MAX_SEG];
    int src_fp32[TL_MAX_SEG];   // the segment's source rows are raw fp32 (converted to the operand format in registers)
};

struct Layout {
    uint32_t b0, idx0, bars, tmem_slot;
    __device__ __forceinline__ uint32_t a_empty(int g, uint32_t r) const { return bars + 8u * (uint32_t)(g * MAX_R + (int)r); }
    __device__ __forceinline__ uint32_t acc_full(int g) const { return bars + 8u * (uint32_t)(4 * MAX_R + g); }
    __device__ __forceinline__ uint32_t b_full(uint32_t s) const { return bars + 8u * (4 * MAX_R + 4 + s); }
    __device__ __forceinline__ uint32_t b_empty(uint32_t s) const { return bars + 8u * (4 * MAX_R + 4 + MAX_NB + s); }
    __device__ __forceinline__ uint32_t wres() const { return bars + 8u * (4 * MAX_R + 4 + 2 * MAX_NB); }
};
constexpr int BAR_BYTES = (8 * (4 * MAX_R + 4 + 2 * MAX_NB + 1) + 15) & ~15;

__device__ __forceinline__ Layout carve(uint32_t base, uint32_t b_bytes, int groups) {
    Layout L;
    L.b0 = base;
    L.idx0 = L.b0 + b_bytes;
    L.bars = L.idx0 + (uint32_t)groups * (IDX_BUF_BYTES + MAX_LIST_BYTES);
    L.tmem_slot = L.bars + BAR_BYTES;
    return L;
}
static inline size_t smem_bytes(size_t b_bytes, int groups) { return 1024 + b_bytes + (size_t)groups * (IDX_BUF_BYTES + MAX_LIST_BYTES) + BAR_BYTES + 32; }

__device__ __forceinline__ uint32_t seg_mask(const tl_conv_seg& sg, int64_t tile) {
    if (!sg.index) return 1u;
    const uint32_t all = sg.n_off >= 32 ? 0xffffffffu : ((1u << sg.n_off) - 1u);
    return (sg.tile_mask ? __ldg(sg.tile_mask + tile) : 0xffffffffu) & all;
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
__device__ __forceinline__ void bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
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
// 16 lanes x 256 bit x 4 (32 columns): r[4g + 0..1] = columns 8g + 2(t%4) + {0,1} of lane t/4, r[4g + 2..3] = the same of lane t/4 + 8
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// One gathered chunk of a thread: 4 rows (lane-quarter rows rr, rr + 8, rr + 16, rr + 24) x 8 channels, hi (and lo) halves
template <int NSPLIT>
struct Frag {
    uint4 hi[4];
    uint4 lo[NSPLIT == 2 ? 4 : 1];
};

// raw fp32 rows (8 floats per lane and row) -> operand format.  The conversion consumes the loads at once, so these
// chunks (the 1x1 projection of the residual stream: two per tile in one conv per level) run in their own unpipelined
// loop; mixed into the pipelined fetch, ptxas predicated both paths on the same instructions and every load of the fp16
// path was waited for in place (profiles/r02_trace_ts_v1_c32.txt: 1000+ cycles per fetch).
template <int NSPLIT>
__device__ __forceinline__ void load_frag_fp32(Frag<NSPLIT>& f, const int (&ix)[4], uint64_t src, uint32_t row_bytes) {
    uint4 u0[4], u1[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        u0[j] = u1[j] = make_uint4(0u, 0u, 0u, 0u);
        if (ix[j] >= 0) {
            const uint64_t p = src + (uint64_t)(uint32_t)ix[j] * row_bytes;
            u0[j] = ldg_nc_v4(p);
            u1[j] = ldg_nc_v4(p + 16);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float v0 = __uint_as_float(u0[j].x), v1 = __uint_as_float(u0[j].y), v2 = __uint_as_float(u0[j].z), v3 = __uint_as_float(u0[j].w);
        const float v4 = __uint_as_float(u1[j].x), v5 = __uint_as_float(u1[j].y), v6 = __uint_as_float(u1[j].z), v7 = __uint_as_float(u1[j].w);
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
// operand-format rows: nothing touches the loaded registers until store_frag
template <int NSPLIT>
__device__ __forceinline__ void load_frag(Frag<NSPLIT>& f, const int (&ix)[4], uint64_t src, uint32_t row_bytes) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint64_t p = src + (uint64_t)(uint32_t)max(ix[j], 0) * row_bytes;
        f.hi[j] = make_uint4(0u, 0u, 0u, 0u);
        if (NSPLIT == 2) f.lo[j] = make_uint4(0u, 0u, 0u, 0u);
        if (ix[j] >= 0) {
            f.hi[j] = ldg_nc_v4(p);
            if (NSPLIT == 2) f.lo[j] = ldg_nc_v4(p + 64);
        }
    }
}
// taddr: TMEM address of the chunk's first column in this warp's lane quarter (lane field = 32 * quarter).
// K step 0 <- bytes [0, 8) of the lane's 16 B piece, K step 1 <- bytes [8, 16).
template <int NSPLIT>
__device__ __forceinline__ void store_frag(const Frag<NSPLIT>& f, uint32_t taddr) {
    tmem_st_16x256b_x2(taddr, f.hi[0].x, f.hi[0].y, f.hi[1].x, f.hi[1].y, f.hi[0].z, f.hi[0].w, f.hi[1].z, f.hi[1].w);
    tmem_st_16x256b_x2(taddr + (16u << 16), f.hi[2].x, f.hi[2].y, f.hi[3].x, f.hi[3].y, f.hi[2].z, f.hi[2].w, f.hi[3].z, f.hi[3].w);
    if (NSPLIT == 2) {
        tmem_st_16x256b_x2(taddr + 16, f.lo[0].x, f.lo[0].y, f.lo[1].x, f.lo[1].y, f.lo[0].z, f.lo[0].w, f.lo[1].z, f.lo[1].w);
        tmem_st_16x256b_x2(taddr + 16 + (16u << 16), f.lo[2].x, f.lo[2].y, f.lo[3].x, f.lo[3].y, f.lo[2].z, f.lo[2].w, f.lo[3].z,
                           f.lo[3].w);
    }
}

template <typename T>
__device__ __forceinline__ T sel3(int s, T a, T b, T c) { return s == 0 ? a : (s == 1 ? b : c); }

// A tile's live chunks as a list in shared memory (built once per tile by the group's first warp), one word per chunk:
//   [31:10] ordinal in the unmasked (segment, offset, k-block) enumeration = weight slab number
//   [9:5] row of the staged rulebook (31 = identity segment)   [4:2] k-block   [1:0] segment
constexpr int MAX_LIST = 448;
constexpr int LIST_BYTES = MAX_LIST * 4 + 16;        // entries + (count, count of operand-format chunks)
constexpr int GROUP_SMEM = IDX_BUF_BYTES + LIST_BYTES;

// Warps 0 .. 4G-1: group g = warp / 4, TMEM lane quarter = warp % 4.  RESIDENT: the layer's weights stay in shared memory
// (loaded once by warp 0).  Otherwise warp 4G streams them through a ring for all groups.
// Tile of (round r, CTA c, group g) = (r * gridDim.x + c) * G + g  (the G tiles a CTA works on at once are neighbours).
template <int NSPLIT, int G, bool RESIDENT>
__global__ void __launch_bounds__(32 * (4 * G + (RESIDENT ? 0 : 1)), 1) k_conv_ts(const tl_conv_desc d, const Launch P) {
    constexpr uint32_t CHUNK_COLS = 16 * NSPLIT;       // TMEM columns of one A chunk
    constexpr int kAuxWarp = RESIDENT ? 0 : 4 * G;     // TMEM allocation (+ the weight stream)
    constexpr int DEPTH = 2;                           // chunk fragments in flight per thread (16 / 32 registers each; 3 spill at G = 4)
    constexpr int FMT = NSPLIT == 2 ? FMT_F16X2 : FMT_F16;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int N = d.c_out;
    const Layout L = carve(base, P.b_bytes, G);
    const uint32_t slab = (uint32_t)N * 64u * NSPLIT;   // weight bytes of one chunk: [C_out][32] fp16 (x hi, lo)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (L.tmem_slot - smem_u32(smem_raw)));
    const uint32_t R = (uint32_t)P.R, NB = (uint32_t)P.nb;
    const int dbg = kTrace ? P.debug : 0;               // timing experiments exist in the TRACE build only

    if (threadIdx.x == 0) {
        for (int g = 0; g < G; ++g) {
            for (uint32_t r = 0; r < R; ++r) mbar_init(L.a_empty(g, r), 1);
            mbar_init(L.acc_full(g), 1);
        }
        for (uint32_t s = 0; s < NB; ++s) {
            mbar_init(L.b_full(s), 1);
            mbar_init(L.b_empty(s), G);
        }
        mbar_init(L.wres(), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kAuxWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(L.tmem_slot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    // weights are packed [n_off][c_in/32][(hi, lo)][C_out][32] (K order + SWIZZLE_64B image applied by
    // sparse.pack_weight_ts): the slabs of a segment are contiguous and a slab lands in its stage as-is
    if (RESIDENT && warp == 0) {
        if (elect_one()) {
            uint32_t total = 0;
            for (int s = 0; s < d.n_seg; ++s) total += (uint32_t)(d.seg[s].n_off * (d.seg[s].c_in / 32)) * slab;
            if (dbg & 8) {
                mbar_arrive(L.wres());
            } else {
                mbar_arrive_expect_tx(L.wres(), total);
                uint32_t dst = L.b0;
                for (int s = 0; s < d.n_seg; ++s) {
                    const uint32_t bytes = (uint32_t)(d.seg[s].n_off * (d.seg[s].c_in / 32)) * slab;
                    const char* src = reinterpret_cast<const char*>(d.seg[s].weight);
                    for (uint32_t off = 0; off < bytes; off += 32768u) bulk_g2s(dst + off, src + off, min(32768u, bytes - off), L.wres());
                    dst += bytes;
                }
            }
        }
        __syncwarp();
    }
    if (!RESIDENT && warp == kAuxWarp) {
        // ===================== weight stream: one for the whole CTA =================================
        // position p = round * chunks_total + ord, slot p % NB, freed by all G groups
        uint32_t nslab[TL_MAX_SEG];
        for (int s = 0; s < TL_MAX_SEG; ++s) nslab[s] = s < d.n_seg ? (uint32_t)(d.seg[s].n_off * (d.seg[s].c_in / 32)) : 0u;
        const uint32_t total = (uint32_t)P.rounds * (uint32_t)P.chunks_total;
        uint32_t slot = 0, phase = 0, ord = 0;
        for (uint32_t p = 0; p < total; ++p) {
            mbar_wait(L.b_empty(slot), phase ^ 1u);
            if (elect_one()) {
                if (dbg & 8) {
                    mbar_arrive(L.b_full(slot));
                } else {
                    const int s = ord < nslab[0] ? 0 : (ord < nslab[0] + nslab[1] ? 1 : 2);
                    const uint32_t local = ord - (s == 0 ? 0u : (s == 1 ? nslab[0] : nslab[0] + nslab[1]));
                    const char* src = reinterpret_cast<const char*>(d.seg[s].weight) + (size_t)local * slab;
                    mbar_arrive_expect_tx(L.b_full(slot), slab);
                    bulk_g2s(L.b0 + slot * slab, src, slab, L.b_full(slot));
                }
            }
            __syncwarp();
            if (++ord == (uint32_t)P.chunks_total) ord = 0;
            if (++slot == NB) slot = 0, phase ^= 1u;
        }
    } else {
        // ===================== a group: gather -> TMEM -> MMA -> epilogue for its own tiles ==========
        const int g = warp >> 2, qtr = warp & 3;
        const int q = lane & 3, rr = lane >> 2;
        const int bar_id = 1 + g;
        const uint32_t lane_field = ((uint32_t)qtr * 32u) << 16;
        const uint32_t acc_col = tmem_base + (uint32_t)(g * P.group_cols);   // accumulator columns of this group
        const uint32_t ring_col = acc_col + (uint32_t)N;                      // its A ring
        const uint32_t idxbuf = L.idx0 + (uint32_t)g * GROUP_SMEM;
        const uint32_t listbuf = idxbuf + IDX_BUF_BYTES;
        const uint32_t idx_thr = idxbuf + (uint32_t)(qtr * 32 + rr) * 4u;     // my first rulebook entry of a staged row
        const bool leader = qtr == 0;
        const uint32_t idesc = make_idesc(N, true);
        const uint64_t bdesc0 = make_smem_desc(L.b0, 64);
        const uint32_t slab16 = slab >> 4, half16 = ((uint32_t)N * 64u) >> 4;
        const int tig = threadIdx.x & 127;                                    // thread in group
        const bool trg = kTrace && (dbg & 32) && blockIdx.x == 0 && g == 0 && lane == 0;
        uint32_t tp = 0;
        // per-segment source: base + this lane's piece, bytes per row (TL_MAX_SEG == 3)
        uint64_t src0 = 0, src1 = 0, src2 = 0;
        uint32_t rb0 = 0, rb1 = 0, rb2 = 0;
        int n_main = d.n_seg;              // segments [0, n_main) are in the operand format, [n_main, n_seg) raw fp32
        for (int s = d.n_seg - 1; s >= 0 && P.src_fp32[s]; --s) n_main = s;
#pragma unroll
        for (int s = 0; s < TL_MAX_SEG; ++s) {
            if (s >= d.n_seg) continue;
            const bool f32 = s >= n_main;
            const uint64_t sb = (uint64_t)d.seg[s].src + (uint32_t)(q * (f32 ? 32 : 16));
            const uint32_t rb = (uint32_t)d.seg[s].src_stride * (f32 ? 4u : 2u * NSPLIT);
            if (s == 0) src0 = sb, rb0 = rb;
            if (s == 1) src1 = sb, rb1 = rb;
            if (s == 2) src2 = sb, rb2 = rb;
        }

        // rulebook rows of a tile -> this group's staging buffer (cp.async, 16 B = 4 tile rows per copy)
        auto stage_index = [&](int tile) {
            if (tile < P.num_tiles) {
#pragma unroll
                for (int s = 0; s < TL_MAX_SEG; ++s) {
                    if (s >= d.n_seg || !d.seg[s].index || !P.idx_owner[s]) continue;
                    const int32_t* ip = d.seg[s].index + (int64_t)tile * BM;
                    const int pieces = d.seg[s].n_off * 32;            // 16 B pieces: 32 per rulebook row
                    for (int e = tig; e < pieces; e += 128) {
                        const int k = e >> 5, c4 = (e & 31) * 4;
                        cp_async16(idxbuf + (uint32_t)((P.idx_base[s] + k) * BM + c4) * 4u, ip + (int64_t)k * d.seg[s].index_stride + c4, 16u);
                    }
                }
            }
            cp_async_commit();
        };
        // the tile's live chunks (first warp of the group; published by the group barrier that follows)
        auto build_list = [&](int tile) {
            const bool valid = tile < P.num_tiles;
            uint32_t pos = 0, ord0 = 0, pos_main = 0;
            for (int s = 0; s < d.n_seg; ++s) {
                const tl_conv_seg& sg = d.seg[s];
                const uint32_t kblocks = (uint32_t)sg.c_in / 32u;
                const uint32_t mask = valid ? seg_mask(sg, tile) : 0u;
                const bool live = lane < sg.n_off && ((mask >> lane) & 1u);
                const uint32_t cnt = live ? kblocks : 0u;
                uint32_t incl = cnt;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, off);
                    if (lane >= off) incl += v;
                }
                const uint32_t irow = sg.index ? (uint32_t)(P.idx_base[s] + lane) : 31u;
                uint32_t at = pos + incl - cnt;
                for (uint32_t kb = 0; kb < cnt; ++kb, ++at)
                    if (at < (uint32_t)MAX_LIST)
                        st_shared_u32(listbuf + 4u * at, ((ord0 + (uint32_t)lane * kblocks + kb) << 10) | (irow << 5) | (kb << 2) | (uint32_t)s);
                pos += __shfl_sync(0xffffffffu, incl, 31);
                ord0 += (uint32_t)sg.n_off * kblocks;
                if (s + 1 == n_main) pos_main = pos;
            }
            if (n_main == 0) pos_main = 0;
            if (lane == 0) {
                st_shared_u32(listbuf + 4u * MAX_LIST, min(pos, (uint32_t)MAX_LIST));
                st_shared_u32(listbuf + 4u * MAX_LIST + 4u, min(pos_main, (uint32_t)MAX_LIST));
            }
        };

        if (RESIDENT) mbar_wait(L.wres(), 0u);
        uint32_t a_slot = 0, a_phase = 0;       // position in this group's A ring
        uint32_t b_slot = 0, b_phase = 0;       // position in the CTA's weight stream (streaming mode; leader warp only)
        stage_index((int)blockIdx.x * G + g);
        if (leader) build_list((int)blockIdx.x * G + g);
        cp_async_wait_all();
        bar_sync(bar_id, 128);

        for (int r = 0; r < P.rounds; ++r) {
            const int tile = (r * (int)gridDim.x + (int)blockIdx.x) * G + g;
            const bool valid = tile < P.num_tiles;
            const int trow = tile * BM + qtr * 32 + rr;          // identity segments: my first row
            const uint32_t n = ld_shared_u32(listbuf + 4u * MAX_LIST), n_op = ld_shared_u32(listbuf + 4u * MAX_LIST + 4u);
            trace(trg, qtr, tp, ((uint32_t)r << 3) | 0u);

            auto rows_of = [&](uint32_t e, int (&ix)[4]) {
                const uint32_t irow = (e >> 5) & 31u;
                if (irow != 31u) {
                    const uint32_t ia = idx_thr + irow * (BM * 4u);
#pragma unroll
                    for (int j = 0; j < 4; ++j) ix[j] = ld_shared_i32(ia + 32u * j);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) ix[j] = (trow + 8 * j < d.n_out) ? trow + 8 * j : -1;
                }
                if (dbg & 2) ix[0] = ix[1] = ix[2] = ix[3] = -1;
            };
            auto fetch = [&](Frag<NSPLIT>& f, uint32_t i) {     // operand-format chunk i of the list
                const uint32_t e = ld_shared_u32(listbuf + 4u * i);
                const int s = (int)(e & 3u);
                int ix[4];
                rows_of(e, ix);
                const uint64_t src = sel3(s, src0, src1, src2) + ((e >> 2) & 7u) * (64u * NSPLIT);
                load_frag<NSPLIT>(f, ix, src, sel3(s, rb0, rb1, rb2));
            };

            // the leader warp walks the weight stream: positions skipped by this tile are released as they are passed
            uint32_t b_next = 0;                // next unconsumed ordinal of this round (streaming mode)
            auto skip_to = [&](uint32_t ord) {  // leader warp, converged
                for (; b_next < ord; ++b_next) {
                    mbar_wait(L.b_full(b_slot), b_phase);
                    if (elect_one()) mbar_arrive(L.b_empty(b_slot));
                    __syncwarp();
                    if (++b_slot == NB) b_slot = 0, b_phase ^= 1u;
                }
            };
            uint32_t first = 0;                 // 0 until the tile's first MMA has been issued (accumulate flag)
            auto process = [&](const Frag<NSPLIT>& f, uint32_t i) {
                mbar_wait(L.a_empty(g, a_slot), a_phase ^ 1u);
                tc_fence_after();
                const uint32_t a_col = ring_col + a_slot * CHUNK_COLS;
                if (!(dbg & 16)) store_frag<NSPLIT>(f, a_col + lane_field);
                tmem_wait_st();
                tc_fence_before();
                bar_sync(bar_id, 128);
                if (leader) {
                    tc_fence_after();
                    const uint32_t ord = ld_shared_u32(listbuf + 4u * i) >> 10;
                    uint32_t boff;
                    if (RESIDENT) {
                        boff = ord * slab16;
                    } else {
                        skip_to(ord);
                        mbar_wait(L.b_full(b_slot), b_phase);
                        boff = b_slot * slab16;
                    }
                    if (elect_one()) {
                        if (!(dbg & 1)) {
                            const uint64_t bd = bdesc0 + (uint64_t)boff;
#pragma unroll
                            for (uint32_t kk = 0; kk < 2; ++kk) {
                                umma_f16_ts(acc_col, a_col + 8 * kk, bd + 2 * kk, idesc, kk | first);
                                if (NSPLIT == 2) {
                                    umma_f16_ts(acc_col, a_col + 8 * kk, bd + half16 + 2 * kk, idesc, 1u);      // hi x lo
                                    umma_f16_ts(acc_col, a_col + 16 + 8 * kk, bd + 2 * kk, idesc, 1u);          // lo x hi
                                }
                            }
                        }
                        umma_commit(L.a_empty(g, a_slot));
                        if (!RESIDENT) umma_commit(L.b_empty(b_slot));
                    }
                    __syncwarp();
                    if (!RESIDENT) {
                        ++b_next;
                        if (++b_slot == NB) b_slot = 0, b_phase ^= 1u;
                    }
                }
                first = 1u;
                if (++a_slot == R) a_slot = 0, a_phase ^= 1u;
            };

            // ---- main loop: chunk fragments are fetched DEPTH-1 chunks ahead of the one being written to TMEM
            Frag<NSPLIT> f[DEPTH];
#pragma unroll
            for (int i = 0; i < DEPTH - 1; ++i)
                if ((uint32_t)i < n_op) fetch(f[i], (uint32_t)i);
            for (uint32_t i0 = 0; i0 < n_op; i0 += DEPTH) {
#pragma unroll
                for (int u = 0; u < DEPTH; ++u) {
                    const uint32_t i = i0 + (uint32_t)u;
                    if (i >= n_op) break;
                    if (i + DEPTH - 1 < n_op) fetch(f[(u + DEPTH - 1) % DEPTH], i + DEPTH - 1);
                    process(f[u], i);
                }
            }
            for (uint32_t i = n_op; i < n; ++i) {      // raw fp32 segments: load, convert, hand over -- one chunk at a time
                const uint32_t e = ld_shared_u32(listbuf + 4u * i);
                const int s = (int)(e & 3u);
                int ix[4];
                rows_of(e, ix);
                // (lane piece of an fp32 row: 32 B; 128 B per 32-channel block)
                load_frag_fp32<NSPLIT>(f[0], ix, sel3(s, src0, src1, src2) + ((e >> 2) & 7u) * 128u, sel3(s, rb0, rb1, rb2));
                process(f[0], i);
            }
            trace(trg, qtr, tp, ((uint32_t)r << 3) | 1u);
            const bool any = n != 0u;
            const int next_tile = ((r + 1) * (int)gridDim.x + (int)blockIdx.x) * G + g;
            if (leader) {
                if (!RESIDENT) skip_to((uint32_t)P.chunks_total);      // release the rest of this round's weight stream
                if (elect_one()) {
                    if (any) umma_commit(L.acc_full(g));
                    else mbar_arrive(L.acc_full(g));
                }
                __syncwarp();
            }
            // every warp of the group is past its last read of this tile's rulebook rows and chunk list (they precede a
            // group barrier): the next tile's stream in behind the epilogue
            stage_index(next_tile);
            if (leader) build_list(next_tile);

            // ---- epilogue: this warp's 32 rows; lane (rr, q) holds positions 8q .. 8q+7 of every 32-channel block of rows
            //      rr, rr + 8 (half 0) and rr + 16, rr + 24 (half 1) of its quarter
            if (valid) {
                const int64_t row0 = (int64_t)tile * BM + qtr * 32 + rr;
                const bool has_res = d.residual != nullptr && !(dbg & 4);
                float4 res[4][2];
                auto fetch_residual = [&](int c0) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int64_t grow = row0 + 8 * i;
                        res[i][0] = res[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (has_res && grow < d.n_out) {
                            const float4* p = reinterpret_cast<const float4*>(d.residual + grow * N + c0 + 8 * q);
                            res[i][0] = __ldg(p);
                            res[i][1] = __ldg(p + 1);
                        }
                    }
                };
                fetch_residual(0);
                mbar_wait(L.acc_full(g), (uint32_t)r & 1u);
                tc_fence_after();
                trace(trg, qtr, tp, ((uint32_t)r << 3) | 2u);
                for (int c0 = 0; c0 < N; c0 += 32) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {          // lanes 16h .. 16h+15 of the quarter: rows rr + 16h and rr + 16h + 8
                        uint32_t v[16];
                        if (any) {
                            tmem_ld_16x256b_x4(acc_col + lane_field + ((uint32_t)(16 * h) << 16) + (uint32_t)c0, v);
                            tmem_wait_ld();
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] = 0u;
                        }
                        float x[2][8];
#pragma unroll
                        for (int ab = 0; ab < 2; ++ab) {
#pragma unroll
                            for (int gg = 0; gg < 4; ++gg) {   // position 2gg + e of my piece = logical channel c0 + 8gg + 2q + e
                                x[ab][2 * gg] = __uint_as_float(v[4 * gg + 2 * ab]);
                                x[ab][2 * gg + 1] = __uint_as_float(v[4 * gg + 2 * ab + 1]);
                            }
                            const float4 r0 = res[2 * h + ab][0], r1 = res[2 * h + ab][1];
                            x[ab][0] += r0.x, x[ab][1] += r0.y, x[ab][2] += r0.z, x[ab][3] += r0.w;
                            x[ab][4] += r1.x, x[ab][5] += r1.y, x[ab][6] += r1.z, x[ab][7] += r1.w;
                        }
                        if (dbg & 4) continue;
                        if (d.out_raw) {
#pragma unroll
                            for (int ab = 0; ab < 2; ++ab) {
                                const int64_t grow = row0 + 16 * h + 8 * ab;
                                if (grow >= d.n_out) continue;
                                float4* p = reinterpret_cast<float4*>(d.out_raw + grow * N + c0 + 8 * q);
                                p[0] = make_float4(x[ab][0], x[ab][1], x[ab][2], x[ab][3]);
                                p[1] = make_float4(x[ab][4], x[ab][5], x[ab][6], x[ab][7]);
                            }
                        }
#pragma unroll
                        for (int w = 0; w < 2; ++w) {
                            void* out = w ? (void*)d.out_act2 : (void*)d.out_act1;
                            if (!out) continue;
                            const float* sp = w ? d.scale2 : d.scale1;
                            const float* tp2 = w ? d.shift2 : d.shift1;
                            float sc[8], sh[8];
#pragma unroll
                            for (int gg = 0; gg < 4; ++gg) {
                                const float2 a = __ldg(reinterpret_cast<const float2*>(sp + c0 + 8 * gg + 2 * q));
                                const float2 b = __ldg(reinterpret_cast<const float2*>(tp2 + c0 + 8 * gg + 2 * q));
                                sc[2 * gg] = a.x, sc[2 * gg + 1] = a.y, sh[2 * gg] = b.x, sh[2 * gg + 1] = b.y;
                            }
#pragma unroll
                            for (int ab = 0; ab < 2; ++ab) {
                                const int64_t grow = row0 + 16 * h + 8 * ab;
                                if (grow >= d.n_out) continue;
                                float a[8];
#pragma unroll
                                for (int j = 0; j < 8; ++j) a[j] = fmaxf(fmaf(x[ab][j], sc[j], sh[j]), 0.f);
                                if (FMT == FMT_F16) {
                                    *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(out) + grow * N + c0 + 8 * q) =
                                        make_uint4(pack_half2(a[0], a[1]), pack_half2(a[2], a[3]), pack_half2(a[4], a[5]), pack_half2(a[6], a[7]));
                                } else {
                                    uint4 hi, lo;
                                    split_half2(a[0], a[1], hi.x, lo.x);
                                    split_half2(a[2], a[3], hi.y, lo.y);
                                    split_half2(a[4], a[5], hi.z, lo.z);
                                    split_half2(a[6], a[7], hi.w, lo.w);
                                    char* p = reinterpret_cast<char*>(out) + (grow * N + c0) * 4 + 16 * q;
                                    *reinterpret_cast<uint4*>(p) = hi;
                                    *reinterpret_cast<uint4*>(p + 64) = lo;
                                }
                            }
                        }
                    }
                    if (c0 + 32 < N) fetch_residual(c0 + 32);
                }
                tc_fence_before();
            } else {
                mbar_wait(L.acc_full(g), (uint32_t)r & 1u);
            }
            trace(trg, qtr, tp, ((uint32_t)r << 3) | 3u);
            cp_async_wait_all();
            bar_sync(bar_id, 128);      // next tile's rulebook rows + chunk list visible; accumulator reads done before its first MMA
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kAuxWarp) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

template <int NSPLIT, int G, bool RESIDENT>
static int launch(const tl_conv_desc& d, const Launch& P, int grid, size_t smem, cudaStream_t stream) {
    static bool configured[16] = {false};
    int dev = 0;
    TL_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 16 && !configured[dev]) {     // function attributes are per device
        TL_CUDA_CHECK(cudaFuncSetAttribute(k_conv_ts<NSPLIT, G, RESIDENT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured[dev] = true;
    }
    k_conv_ts<NSPLIT, G, RESIDENT><<<grid, 32 * (4 * G + (RESIDENT ? 0 : 1)), smem, stream>>>(d, P);
    TL_LAUNCH_CHECK();
    return TL_OK;
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
    return idx_rows <= ts::IDX_ROWS && chunks > 0 && chunks <= 448;
}

// nsplit: 1 = TL_MODE_F16, 2 = TL_MODE_F16X2.  src_fp32_mask bit s: segment s reads raw fp32 rows.
int conv_fwd_ts(const tl_conv_desc& d, cudaStream_t stream, int nsplit, uint32_t src_fp32_mask) {
    static int num_sms[16] = {0};
    int dev = 0;
    TL_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 16) {
        set_error("tl_conv_fwd: device ordinal %d not supported", dev);
        return TL_ERR_UNSUPPORTED;
    }
    if (!num_sms[dev]) TL_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms[dev], cudaDevAttrMultiProcessorCount, dev));
    const int sms = num_sms[dev];
    const int n = d.c_out;
    for (int s = 0; s + 1 < d.n_seg; ++s)
        if (((src_fp32_mask >> s) & 1u) && !((src_fp32_mask >> (s + 1)) & 1u)) {
            set_error("tl_conv_fwd(ts): raw fp32 segments must come after the operand-format segments");
            return TL_ERR_ARG;
        }
    ts::Launch P;
    memset(&P, 0, sizeof(P));
    P.debug = env_int_ts("TL_TS_DEBUG", 0);
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
    // groups per CTA: as many as tensor memory allows with >= 2 A chunks per group ring (4 groups at C_out <= 96 fp16)
    const int chunk_cols = 16 * nsplit;
    int groups = 0;
    const int cand[4] = {4, 3, 2, 1};
    for (int i = 0; i < 4 && !groups; ++i) {
        const int cols = (ts::TMEM_COLS / cand[i]) & ~15;
        if (cols - n >= 2 * chunk_cols) groups = cand[i];
    }
    if (env_int_ts("TL_TS_GROUPS", 0) > 0 && env_int_ts("TL_TS_GROUPS", 0) < groups) groups = env_int_ts("TL_TS_GROUPS", 0);
    if (!groups) {
        set_error("tl_conv_fwd(ts): c_out=%d leaves no room for the TMEM A ring", n);
        return TL_ERR_UNSUPPORTED;
    }
    P.group_cols = (ts::TMEM_COLS / groups) & ~15;
    P.R = (P.group_cols - n) / chunk_cols;
    if (P.R > ts::MAX_R) P.R = ts::MAX_R;
    if (env_int_ts("TL_TS_R", 0) >= 2 && env_int_ts("TL_TS_R", 0) < P.R) P.R = env_int_ts("TL_TS_R", 0);
    // weights: resident when the whole layer fits beside the rulebook staging, else a ring of slabs
    const uint32_t slab = (uint32_t)n * 64u * nsplit;
    const size_t fixed = ts::smem_bytes(0, groups);
    const size_t budget = (size_t)env_int_ts("TL_TS_SMEM_KB", 200) * 1024;
    P.resident = (fixed + (size_t)P.chunks_total * slab <= budget) ? 1 : 0;
    if (env_int_ts("TL_TS_RESIDENT", 1) == 0) P.resident = 0;
    if (P.resident) {
        P.nb = 0;
        P.b_bytes = (uint32_t)P.chunks_total * slab;
    } else {
        int nb = (int)((budget - fixed) / slab);
        if (nb > ts::MAX_NB) nb = ts::MAX_NB;
        if (nb < 2) {
            set_error("tl_conv_fwd(ts): weight slab of %u bytes does not fit the shared-memory ring", slab);
            return TL_ERR_UNSUPPORTED;
        }
        P.nb = nb;
        P.b_bytes = (uint32_t)nb * slab;
    }
    const size_t smem = ts::smem_bytes(P.b_bytes, groups);
    int grid = (P.num_tiles + groups - 1) / groups;
    if (grid > sms) grid = sms;
    P.rounds = (P.num_tiles + grid * groups - 1) / (grid * groups);
#define TL_TS_LAUNCH(NS, GG) return P.resident ? ts::launch<NS, GG, true>(d, P, grid, smem, stream) : ts::launch<NS, GG, false>(d, P, grid, smem, stream)
    if (nsplit == 1) {
        switch (groups) {
            case 4: TL_TS_LAUNCH(1, 4);
            case 3: TL_TS_LAUNCH(1, 3);
            case 2: TL_TS_LAUNCH(1, 2);
            default: TL_TS_LAUNCH(1, 1);
        }
    }
    switch (groups) {
        case 4: TL_TS_LAUNCH(2, 4);
        case 3: TL_TS_LAUNCH(2, 3);
        case 2: TL_TS_LAUNCH(2, 2);
        default: TL_TS_LAUNCH(2, 1);
    }
#undef TL_TS_LAUNCH
}

}  // namespace tl

// debug: copy the timeline trace of the last TL_TS_DEBUG=32 launch to the host (roles x TRACE_LEN u64 = tag << 48 | clock)
extern "C" int tl_debug_copy_trace_ts(void* host, size_t bytes) {
    const size_t want = sizeof(unsigned long long) * tl::ts::TRACE_ROLES * tl::ts::TRACE_LEN;
    if (bytes < want) return TL_ERR_ARG;
    cudaDeviceSynchronize();
    return cudaMemcpyFromSymbol(host, tl::ts::g_trace_ts, want) == cudaSuccess ? TL_OK : TL_ERR_CUDA;
}
