// tcgen05 sparse convolution, "group" form, sm_100a only.  Modes TL_MODE_F16 (fp16 operands) and TL_MODE_F16X2 (two-term
// fp16 split of both operands, three MMAs per K step = fp32-equivalent products); fp32 accumulation in tensor memory.
//
// What round 2 measured before this kernel (profiles/r02_conv_history.md):
//   * round 1's kernel (warp roles: 12 gather warps -> 1 MMA warp -> 4 epilogue warps, a scheduler warp) is bound by the
//     mbarrier hand-offs between the roles: its empty skeleton costs 4 500 cycles per 128-row tile;
//   * moving the gathered A operand through registers into tensor memory (TS-form MMA: 16.6 instead of 41 cycles per
//     N = 32 MMA) makes the kernel INSTRUCTION-bound: ~460 -> 150 issued instructions per chunk and warp (index loads,
//     address arithmetic, zero fill, the register shuffle tcgen05.st wants), 2-3x more than a cp.async gather needs.
// Hence: cp.async gathers (no registers, ~35 instructions per chunk and thread), the shared-memory (SS) MMA, and NO
// cross-role hand-off on the critical path.  The CTA (one per SM, persistent) is split into G independent GROUPS of four
// warps.  A group owns one 128-row tile at a time, a private accumulator in tensor memory and a private ring of R A
// stages in shared memory, and does everything for its tile:
//   * gather (warps 1-3 of the group, chunk i of the tile by warp 1 + i % 3): ONE warp copies all 128 neighbour rows of a
//     chunk (a chunk = 128 rows x 32 channels of one (segment, offset, k-block)) -- lane (piece c = lane & 3, sub-row lane >> 2)
//     the 16 B piece c of rows sub + 8 i, i = 0..15 -- with cp.async.cg into the SWIZZLE_64B K-major stage, so the per-chunk
//     bookkeeping is paid once per 16 copies (with 4 copies per thread the kernel was instruction-bound, profiles/
//     r02_conv_history.md); the rulebook entries come straight from global memory one chunk ahead; absent neighbours read one
//     of 256 spread-out zero rows (uniform 16 B copies); completion = cp.async.mbarrier.arrive.noinc on the stage's `full`
//     barrier -- nobody waits for its own copies;
//   * three chunks (one per gather warp) form a FILL with one `full` / `empty` mbarrier pair: the group's first warp waits for
//     `full` of the fills in order and its elected lane issues the fill's tcgen05.mma K steps + ONE tcgen05.commit -> `empty`
//     (with a barrier pair per chunk this warp was the bottleneck: ~550 cycles per chunk for 82 cycles of MMA,
//     profiles/r02_trace_grp_v5_c32_chunks.txt);
//   * epilogue: the group's four warps are the four TMEM lane quarters; tcgen05.ld.16x256b hands every lane 8 channels of
//     4 rows, written as 16 B / 32 B vectors that a quad of lanes makes a contiguous row piece: no shared memory.
// Weights: resident in shared memory when the layer fits (C = 32: 54 KB), else one stream per CTA (TMA bulk copies into
// a ring, warp 4G) feeds all G groups, which walk the kernel offsets in the same order.
//
// Channel order ("P-layout", treelearn_b200/sparse.py): within every 32-channel block the tensors this kernel reads and
// writes (fp32 residual stream, activated operands) store logical channel 8g + 2q + e at position 8q + 2g + e; that makes
// the 8 accumulator columns a lane receives from tcgen05.ld.16x256b contiguous in memory.  K order of the weights =
// memory order of the rows (sparse.pack_weight_grp); scale / shift vectors stay in logical order.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "tl_common.cuh"
#include "tl_tc_ptx.cuh"

namespace tl {
namespace grp {

using namespace tl::tc;

constexpr int MAX_G = 4;
constexpr int MAX_R = 9;                   // A stages per group (a multiple of GW)
constexpr int MAX_NB = 64;                 // weight ring slots
constexpr int MAX_LIST = 448;              // live chunks of a tile
constexpr int LIST_BYTES = MAX_LIST * 4 + 16;
constexpr int GROUP_SMEM = LIST_BYTES;     // per group, beside its A ring: the tile's chunk list
constexpr int GW = 3;                      // gather warps per group; chunk i of a tile is gathered by warp 1 + i % GW
constexpr int TMEM_COLS = 512;

// 256 zero regions of 1 KB: source of absent neighbour rows (never written).  Spread out so the reads do not serialise on
// one L2 line (round 1: a single shared zero row made the gather 8x slower).
__device__ float g_zero_rows_grp[256 * 256];

// ---- optional timeline trace (make TRACE=1, TL_GRP_DEBUG bit 32): group 0 of CTA 0 records clock64() (tools/trace_grp.py)
constexpr int TRACE_ROLES = 8, TRACE_LEN = 4096;
__device__ unsigned long long g_trace_grp[TRACE_ROLES * TRACE_LEN];
#ifdef TL_TC_TRACE
constexpr bool kTrace = true;
#else
constexpr bool kTrace = false;
#endif
__device__ __forceinline__ void trace(bool on, int role, uint32_t& pos, uint32_t tag) {
    if (kTrace && on && pos < TRACE_LEN) {
        g_trace_grp[role * TRACE_LEN + pos] = ((unsigned long long)tag << 48) | ((unsigned long long)clock64() & 0xffffffffffffull);
        ++pos;
    }
}

struct Launch {
    int num_tiles, rounds, chunks_total;
    int R, D;          // A stages per group; MMA issue runs D chunks behind the gather front (D < R)
    int acc_cols;      // TMEM columns per group
    int resident;      // 1: every weight slab of the layer stays in shared memory; 0: weight ring of `nb` slabs
    int nb;
    int fill;          // chunks per full/empty barrier pair: 1, or GW (one chunk per gather warp)
    uint32_t b_bytes;  // shared-memory bytes of the weight region
    int debug;         // TRACE build: 1 skip MMAs, 2 skip row copies, 4 skip epilogue memory ops, 8 skip weight copies
};

struct Layout {
    uint32_t b0, a0, grp0, bars, tmem_slot, a_group_bytes;
    __device__ __forceinline__ uint32_t full(int g, uint32_t r) const { return bars + 8u * (uint32_t)(g * MAX_R + (int)r); }
    __device__ __forceinline__ uint32_t empty(int g, uint32_t r) const { return bars + 8u * (uint32_t)(MAX_G * MAX_R + g * MAX_R + (int)r); }
    __device__ __forceinline__ uint32_t acc_full(int g) const { return bars + 8u * (uint32_t)(2 * MAX_G * MAX_R + g); }
    __device__ __forceinline__ uint32_t b_full(uint32_t s) const { return bars + 8u * (2 * MAX_G * MAX_R + MAX_G + s); }
    __device__ __forceinline__ uint32_t b_empty(uint32_t s) const { return bars + 8u * (2 * MAX_G * MAX_R + MAX_G + MAX_NB + s); }
    __device__ __forceinline__ uint32_t wres() const { return bars + 8u * (2 * MAX_G * MAX_R + MAX_G + 2 * MAX_NB); }
};
constexpr int BAR_BYTES = (8 * (2 * MAX_G * MAX_R + MAX_G + 2 * MAX_NB + 1) + 15) & ~15;

// [weights][A rings: G x R stages, 1 KB aligned][per group: chunk list][barriers]
__device__ __forceinline__ Layout carve(uint32_t base, uint32_t b_bytes, int groups, int R, uint32_t stage_bytes) {
    Layout L;
    L.b0 = base;
    L.a0 = (L.b0 + b_bytes + 1023u) & ~1023u;
    L.a_group_bytes = (uint32_t)R * stage_bytes;
    L.grp0 = L.a0 + (uint32_t)groups * L.a_group_bytes;
    L.bars = L.grp0 + (uint32_t)groups * GROUP_SMEM;
    L.tmem_slot = L.bars + BAR_BYTES;
    return L;
}
static inline size_t smem_bytes(size_t b_bytes, int groups, int R, size_t stage_bytes) {
    return 1024 + ((b_bytes + 1023) & ~(size_t)1023) + (size_t)groups * ((size_t)R * stage_bytes + GROUP_SMEM) + BAR_BYTES + 32;
}

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
__device__ __forceinline__ void bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
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
template <typename T>
__device__ __forceinline__ T sel3(int s, T a, T b, T c) { return s == 0 ? a : (s == 1 ? b : c); }

// A tile's live chunks as a list in shared memory (built once per tile by the group's first warp), one word per chunk:
//   [31:10] ordinal in the unmasked (segment, offset, k-block) enumeration = weight slab number
//   [9:5] kernel offset   [4:2] k-block   [1:0] segment
//
// Warps 0 .. 4G-1: group g = warp / 4, TMEM lane quarter = warp % 4.  RESIDENT: the layer's weights stay in shared memory
// (loaded once by warp 0).  Otherwise warp 4G streams them through a ring for all groups.
// Tile of (round r, CTA c, group g) = (r * gridDim.x + c) * G + g  (the G tiles a CTA works on at once are neighbours).
template <int NSPLIT, int G, bool RESIDENT>
__global__ void __launch_bounds__(32 * (4 * G + (RESIDENT ? 0 : 1)), 1) k_conv_grp(const tl_conv_desc d, const Launch P) {
    constexpr uint32_t STAGE = (uint32_t)BM * 64u * NSPLIT;   // bytes of one A stage: 128 rows x 64 B (x hi, lo)
    constexpr int kAuxWarp = RESIDENT ? 0 : 4 * G;            // TMEM allocation (+ the weight stream)
    constexpr int FMT = NSPLIT == 2 ? FMT_F16X2 : FMT_F16;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int N = d.c_out;
    const uint32_t R = (uint32_t)P.R, NB = (uint32_t)P.nb;
    const uint32_t FILL = (uint32_t)P.fill, F = R / FILL;      // chunks per fill (1 or GW), fill slots of a group's ring
    const Layout L = carve(base, P.b_bytes, G, P.R, STAGE);
    const uint32_t slab = (uint32_t)N * 64u * NSPLIT;   // weight bytes of one chunk: [C_out][32] fp16 (x hi, lo)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (L.tmem_slot - smem_u32(smem_raw)));
    const int dbg = kTrace ? P.debug : 0;               // timing experiments exist in the TRACE build only

    if (threadIdx.x == 0) {
        for (int g = 0; g < G; ++g) {
            for (uint32_t f = 0; f < R / (uint32_t)P.fill; ++f) {     // per fill = 1 stage, or GW stages (one per gather warp)
                mbar_init(L.full(g, f), 32 * P.fill);    // every lane of the fill's gather warps: its copies have landed
                mbar_init(L.empty(g, f), 1);            // tcgen05.commit of the MMAs that read the fill
            }
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

    // weights are packed [n_off][c_in/32][(hi, lo)][C_out][32] (P-layout K order + SWIZZLE_64B image applied by
    // sparse.pack_weight_grp): the slabs of a segment are contiguous and a slab lands in its stage as-is
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
        // ===================== a group: gather -> MMA -> epilogue for its own tiles =================
        const int g = warp >> 2, qtr = warp & 3;
        const int q = lane & 3, rr = lane >> 2;          // epilogue: piece q of rows rr + 8 i of my lane quarter
        const int pc = lane & 3, sub = lane >> 2;         // gather: 16 B piece pc of rows sub + 8 i, i = 0..15
        const int gw = qtr - 1;                           // gather warps 0..2 of the group (warp 0 issues the MMAs)
        const int bar_id = 1 + g;
        const uint32_t lane_field = ((uint32_t)qtr * 32u) << 16;
        const uint32_t acc_col = tmem_base + (uint32_t)(g * P.acc_cols);
        const uint32_t ring = L.a0 + (uint32_t)g * L.a_group_bytes;
        const uint32_t listbuf = L.grp0 + (uint32_t)g * GROUP_SMEM;
        // 16 B piece c of row r lives at c ^ ((r >> 1) & 3) (SWIZZLE_64B); rows sub + 8 i share the swizzle term
        const uint32_t dst_thr = (uint32_t)(sub * 64 + ((pc ^ ((sub >> 1) & 3)) << 4));
        const bool leader = qtr == 0;
        const uint32_t idesc = make_idesc(N, true);
        const uint64_t bdesc0 = make_smem_desc(L.b0, 64);
        const uint64_t adesc0 = make_smem_desc(ring, 64);
        const uint32_t slab16 = slab >> 4, half16 = ((uint32_t)N * 64u) >> 4, stage16 = STAGE >> 4, ahalf16 = ((uint32_t)BM * 64u) >> 4;
        const uint64_t zero_src = (uint64_t)g_zero_rows_grp + (uint32_t)(pc * 16);
        const bool trg = kTrace && (dbg & 32) && blockIdx.x == 0 && g == 0 && lane == 0;
        const bool trc = trg && qtr == 1, trl = trg && qtr == 0;      // per-chunk events of gather warp 1 / the MMA warp
        uint32_t tp = 0, tp2 = 0;
        // per-segment source: base + this thread's piece, bytes per row (TL_MAX_SEG == 3)
        uint64_t src0 = 0, src1 = 0, src2 = 0;
        uint32_t rb0 = 0, rb1 = 0, rb2 = 0;
        const int32_t *ip0 = nullptr, *ip1 = nullptr, *ip2 = nullptr;    // rulebook of the segment + my sub-row (null: identity)
        uint32_t is0 = 0, is1 = 0, is2 = 0;                              // its row stride (entries)
#pragma unroll
        for (int s = 0; s < TL_MAX_SEG; ++s) {
            if (s >= d.n_seg) continue;
            const uint64_t sb = (uint64_t)d.seg[s].src + (uint32_t)(pc * 16);
            const uint32_t rb = (uint32_t)d.seg[s].src_stride * (2u * NSPLIT);
            const int32_t* ip = d.seg[s].index ? d.seg[s].index + sub : nullptr;
            const uint32_t is = (uint32_t)d.seg[s].index_stride;
            if (s == 0) src0 = sb, rb0 = rb, ip0 = ip, is0 = is;
            if (s == 1) src1 = sb, rb1 = rb, ip1 = ip, is1 = is;
            if (s == 2) src2 = sb, rb2 = rb, ip2 = ip, is2 = is;
        }

        // the tile's live chunks (first warp of the group; published by the group barrier that follows)
        auto build_list = [&](int tile) {
            const bool valid = tile < P.num_tiles;
            uint32_t pos = 0, ord0 = 0;
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
                uint32_t at = pos + incl - cnt;
                for (uint32_t kb = 0; kb < cnt; ++kb, ++at)
                    if (at < (uint32_t)MAX_LIST)
                        st_shared_u32(listbuf + 4u * at, ((ord0 + (uint32_t)lane * kblocks + kb) << 10) | ((uint32_t)lane << 5) | (kb << 2) | (uint32_t)s);
                pos += __shfl_sync(0xffffffffu, incl, 31);
                ord0 += (uint32_t)sg.n_off * kblocks;
            }
            if (lane == 0) st_shared_u32(listbuf + 4u * MAX_LIST, min(pos, (uint32_t)MAX_LIST));
        };

        if (RESIDENT) mbar_wait(L.wres(), 0u);
        uint32_t g_slot = 0, g_phase = 0;       // ring position (in fills) of the tile's first fill
        uint32_t m_slot = 0, m_phase = 0;       // MMA front (first warp), in fills
        uint32_t b_slot = 0, b_phase = 0;       // position in the CTA's weight stream (streaming mode; first warp)
        if (leader) build_list((int)blockIdx.x * G + g);
        bar_sync(bar_id, 128);

        for (int r = 0; r < P.rounds; ++r) {
            const int tile = (r * (int)gridDim.x + (int)blockIdx.x) * G + g;
            const bool valid = tile < P.num_tiles;
            const uint32_t n = ld_shared_u32(listbuf + 4u * MAX_LIST);
            const uint32_t zr0 = (uint32_t)tile * 37u + (uint32_t)sub * 32u;
            trace(trg, qtr, tp, ((uint32_t)r << 3) | 0u);

            // ---- rulebook entries of chunk i (rows sub + 8 j of the tile, j = 0..15) -> registers, one chunk ahead of its gather
            auto load_idx = [&](uint32_t i, int (&ix)[16]) {
                const uint32_t e = ld_shared_u32(listbuf + 4u * i);
                const int s = (int)(e & 3u);
                const int32_t* ip = sel3(s, ip0, ip1, ip2);
                if (ip) {
                    ip += (size_t)((e >> 5) & 31u) * sel3(s, is0, is1, is2) + (size_t)tile * BM;
#pragma unroll
                    for (int j = 0; j < 16; ++j) ix[j] = __ldg(ip + 8 * j);
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) ix[j] = (tile * BM + sub + 8 * j < d.n_out) ? tile * BM + sub + 8 * j : -1;
                }
            };
            // ---- one warp copies all 128 rows of chunk i into A stage `stage` of the group's ring
            auto gather = [&](uint32_t i, const int (&ix)[16], uint32_t stage) {
                const uint32_t e = ld_shared_u32(listbuf + 4u * i);
                const int s = (int)(e & 3u);
                const uint32_t kboff = ((e >> 2) & 7u) * (64u * NSPLIT);
                const uint64_t src = sel3(s, src0, src1, src2) + kboff;
                const uint32_t rb = sel3(s, rb0, rb1, rb2);
                const uint32_t dst = ring + stage * STAGE + dst_thr;
                if (!(dbg & 2)) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        // absent neighbours read one of 256 spread-out zero rows: every copy is a uniform 16 B
                        const uint64_t z = zero_src + (((zr0 + (uint32_t)j) & 255u) << 10);
                        const uint64_t p = ix[j] >= 0 ? src + (uint64_t)(uint32_t)ix[j] * rb : z;
                        cp_async16_cg(dst + (uint32_t)(j * 8 * 64), reinterpret_cast<const void*>(p), 16u);
                        if (NSPLIT == 2)
                            cp_async16_cg(dst + (uint32_t)(BM * 64 + j * 8 * 64), reinterpret_cast<const void*>(ix[j] >= 0 ? p + 64 : z), 16u);
                    }
                }
            };
            // ---- first warp: the MMAs of fill f (chunks GW f .. GW f + cnt - 1; stage GW * slot + c), one barrier round trip
            uint32_t b_next = 0;                // next unconsumed ordinal of this round (streaming mode)
            auto skip_to = [&](uint32_t ord) {  // first warp, converged
                for (; b_next < ord; ++b_next) {
                    mbar_wait(L.b_full(b_slot), b_phase);
                    if (elect_one()) mbar_arrive(L.b_empty(b_slot));
                    __syncwarp();
                    if (++b_slot == NB) b_slot = 0, b_phase ^= 1u;
                }
            };
            auto issue_fill = [&](uint32_t f) {
                const uint32_t i0 = FILL * f, cnt = min(FILL, n - i0);
                trace(trl, 4, tp2, (f << 3) | 0u);
                mbar_wait(L.full(g, m_slot), m_phase);
                trace(trl, 4, tp2, (f << 3) | 1u);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // cp.async (generic proxy) writes -> tensor core reads
                tc_fence_after();
#pragma unroll
                for (uint32_t c = 0; c < (uint32_t)GW; ++c) {
                    if (c >= cnt) break;
                    // the weight slab is taken only now: holding ring slots across the A wait can deadlock when the
                    // tile's list skips more ordinals than the ring has slots
                    const uint32_t ord = ld_shared_u32(listbuf + 4u * (i0 + c)) >> 10;
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
                            const uint64_t ad = adesc0 + (uint64_t)((m_slot * FILL + c) * stage16);
                            const uint64_t bd = bdesc0 + (uint64_t)boff;
#pragma unroll
                            for (uint32_t kk = 0; kk < 2; ++kk) {
                                umma_f16(acc_col, ad + 2 * kk, bd + 2 * kk, idesc, (i0 | c | kk) ? 1u : 0u);
                                if (NSPLIT == 2) {
                                    umma_f16(acc_col, ad + 2 * kk, bd + half16 + 2 * kk, idesc, 1u);            // hi x lo
                                    umma_f16(acc_col, ad + ahalf16 + 2 * kk, bd + 2 * kk, idesc, 1u);           // lo x hi
                                }
                            }
                        }
                        if (!RESIDENT) umma_commit(L.b_empty(b_slot));
                        if (c + 1 == cnt) umma_commit(L.empty(g, m_slot));
                    }
                    __syncwarp();
                    if (!RESIDENT) {
                        ++b_next;
                        if (++b_slot == NB) b_slot = 0, b_phase ^= 1u;
                    }
                }
                trace(trl, 4, tp2, (f << 3) | 2u);
                if (++m_slot == F) m_slot = 0, m_phase ^= 1u;
            };

            const uint32_t nf = FILL == 1 ? n : (n + GW - 1) / GW;
            if (!leader) {
                // FILL == GW: gather warp gw takes chunk GW f + gw of every fill f (stage GW slot + gw).
                // FILL == 1:  every chunk is its own fill; warp gw takes fills gw, gw + GW, ...
                // The ring position of fill f is (g_slot + f) mod F either way.
                const uint32_t f0 = FILL == 1 ? (uint32_t)gw : 0u, fstep = FILL == 1 ? (uint32_t)GW : 1u;
                const uint32_t cfill = FILL == 1 ? 0u : (uint32_t)gw;
                uint32_t slot = g_slot + f0, phase = g_phase;
                if (slot >= F) slot -= F, phase ^= 1u;
                int ixa[16], ixb[16];
                if ((uint32_t)gw < n) load_idx((uint32_t)gw, ixa);
                for (uint32_t f = f0; f < nf; f += 2 * fstep) {
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const uint32_t fu = f + (uint32_t)u * fstep;
                        if (fu >= nf) break;
                        const uint32_t i = FILL * fu + cfill;             // == gw (mod GW) in both schemes
                        if (i + GW < n) load_idx(i + GW, u ? ixa : ixb);
                        trace(trc, 5, tp2, (i << 3) | 0u);
                        mbar_wait(L.empty(g, slot), phase ^ 1u);
                        trace(trc, 5, tp2, (i << 3) | 1u);
                        if (i < n) {
                            gather(i, u ? ixb : ixa, slot * FILL + cfill);
                            cp_async_mbar_arrive_noinc(L.full(g, slot));
                        } else {
                            mbar_arrive(L.full(g, slot));          // a short last fill: nothing to copy for this warp
                        }
                        trace(trc, 5, tp2, (i << 3) | 2u);
                        slot += fstep;
                        if (slot >= F) slot -= F, phase ^= 1u;
                    }
                }
            } else {
                for (uint32_t f = 0; f < nf; ++f) issue_fill(f);
            }
            // ring position of the next tile's first fill
            for (uint32_t left = nf; left;) {     // (no integer division: a tile has a few dozen fills at most)
                const uint32_t step = min(left, F - g_slot);
                g_slot += step, left -= step;
                if (g_slot == F) g_slot = 0, g_phase ^= 1u;
            }
            if (leader) {
                if (!RESIDENT) skip_to((uint32_t)P.chunks_total);      // release the rest of this round's weight stream
                if (elect_one()) {
                    if (n) umma_commit(L.acc_full(g));
                    else mbar_arrive(L.acc_full(g));
                }
                __syncwarp();
            }
            trace(trg, qtr, tp, ((uint32_t)r << 3) | 1u);
            const bool any = n != 0u;
            const int next_tile = ((r + 1) * (int)gridDim.x + (int)blockIdx.x) * G + g;
            // all four warps must be past their last read of this tile's chunk list before the next one replaces it
            bar_sync(bar_id, 128);
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
            bar_sync(bar_id, 128);      // next tile's chunk list visible; accumulator reads done before its first MMA
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
        TL_CUDA_CHECK(cudaFuncSetAttribute(k_conv_grp<NSPLIT, G, RESIDENT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured[dev] = true;
    }
    k_conv_grp<NSPLIT, G, RESIDENT><<<grid, 32 * (4 * G + (RESIDENT ? 0 : 1)), smem, stream>>>(d, P);
    TL_LAUNCH_CHECK();
    return TL_OK;
}

}  // namespace grp

static int env_int_grp(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

bool conv_grp_eligible(const tl_conv_desc& d, uint32_t src_fp32_mask) {
    if (d.c_out % 32 != 0 || d.c_out > 256 || src_fp32_mask) return false;
    int chunks = 0;
    for (int s = 0; s < d.n_seg; ++s) {
        const tl_conv_seg& g = d.seg[s];
        if (g.c_in % 32 != 0 || g.c_in > 256 || g.src_stride % 8 != 0) return false;
        chunks += g.n_off * (g.c_in / 32);
    }
    return chunks > 0 && chunks <= grp::MAX_LIST;
}

// nsplit: 1 = TL_MODE_F16, 2 = TL_MODE_F16X2
int conv_fwd_grp(const tl_conv_desc& d, cudaStream_t stream, int nsplit) {
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
    grp::Launch P;
    memset(&P, 0, sizeof(P));
    P.debug = env_int_grp("TL_GRP_DEBUG", 0);
    P.num_tiles = (d.n_out + tc::BM - 1) / tc::BM;
    P.chunks_total = 0;
    for (int s = 0; s < d.n_seg; ++s) P.chunks_total += d.seg[s].n_off * (d.seg[s].c_in / 32);
    const size_t stage = (size_t)tc::BM * 64 * nsplit;
    const uint32_t slab = (uint32_t)n * 64u * nsplit;
    const size_t budget = (size_t)env_int_grp("TL_GRP_SMEM_KB", 224) * 1024;
    const size_t wbytes = (size_t)P.chunks_total * slab;
    // Chunks per barrier round trip: one for the narrow layers (most chunks in flight matters: the loaded L2 latency is
    // ~5000 cycles, profiles/r02_trace_grp_v7_c32.txt), GW for c_out >= 96 (f16) / >= 160 (f16x2) where the MMA warp's per-chunk
    // hand-off is the bottleneck (profiles/r02_layers_cfg2_f16_grp_v7.txt, r02_grp_v7_fill_f16x2.txt)
    int fill = (nsplit == 1 ? n >= 96 : n >= 160) ? grp::GW : 1;
    if (env_int_grp("TL_GRP_FILL", 0) == 1 || env_int_grp("TL_GRP_FILL", 0) == grp::GW) fill = env_int_grp("TL_GRP_FILL", 0);
    P.fill = fill;
    const int minR = fill == 1 ? 3 : 2 * fill, maxR = grp::MAX_R / fill * fill;
    // groups x stages: resident weights if they leave enough room, else a weight ring of >= 4 slabs
    int bestG = 0, bestR = 0, best_res = 0, best_score = -1;
    for (int res = 1; res >= 0; --res) {
        if (res && env_int_grp("TL_GRP_RESIDENT", 1) == 0) continue;
        const size_t bmin = res ? wbytes : (size_t)4 * slab;
        for (int G = grp::MAX_G; G >= 1; --G) {
            if (G * n > grp::TMEM_COLS) continue;
            if (G > 1 && (G - 1) * sms >= P.num_tiles) continue;     // small levels: spread the tiles over the SMs first
            for (int R = maxR; R >= minR; R -= fill) {
                if (grp::smem_bytes(bmin, G, R, stage) > budget) continue;
                // fill 1: stages beyond 5 per group add nothing (3 gather warps + MMA); fill GW: whole fills, up to 3
                const int score = fill == 1 ? G * (R - 1 < 4 ? R - 1 : 4) : G * (R / fill < 3 ? R / fill : 3);
                if (score > best_score) best_score = score, bestG = G, bestR = R, best_res = res;
                break;
            }
        }
        if (best_score >= (fill == 1 ? 9 : 6)) break;     // resident weights with enough gather depth: take it
    }
    if (env_int_grp("TL_GRP_GROUPS", 0) > 0) {      // experiments: force G, largest R that fits
        bestG = env_int_grp("TL_GRP_GROUPS", 0);
        bestR = 0;
        for (int R = maxR; R >= minR && !bestR; R -= fill)
            if (grp::smem_bytes(best_res ? wbytes : (size_t)4 * slab, bestG, R, stage) <= budget) bestR = R;
    }
    if (env_int_grp("TL_GRP_R", 0) >= minR && env_int_grp("TL_GRP_R", 0) <= bestR) bestR = env_int_grp("TL_GRP_R", 0) / fill * fill;
    if (bestG < 1 || bestR < minR || bestG * n > grp::TMEM_COLS) {
        set_error("tl_conv_fwd(grp): no shared-memory configuration for c_out=%d, %d chunks", n, P.chunks_total);
        return TL_ERR_UNSUPPORTED;
    }
    const int G = bestG;
    P.R = bestR;
    P.D = 0;
    P.acc_cols = n;
    P.resident = best_res;
    if (P.resident) {
        P.nb = 0;
        P.b_bytes = (uint32_t)wbytes;
    } else {
        size_t room = budget - grp::smem_bytes(0, G, P.R, stage);
        int nb = (int)(room / slab);
        if (nb > grp::MAX_NB) nb = grp::MAX_NB;
        if (nb < 2) {
            set_error("tl_conv_fwd(grp): weight slab of %u bytes does not fit the shared-memory ring", slab);
            return TL_ERR_UNSUPPORTED;
        }
        P.nb = nb;
        P.b_bytes = (uint32_t)nb * slab;
    }
    const size_t smem = grp::smem_bytes(P.b_bytes, G, P.R, stage);
    int grid = (P.num_tiles + G - 1) / G;
    if (grid > sms) grid = sms;
    P.rounds = (P.num_tiles + grid * G - 1) / (grid * G);
#define TL_GRP_LAUNCH(NS, GG) return P.resident ? grp::launch<NS, GG, true>(d, P, grid, smem, stream) : grp::launch<NS, GG, false>(d, P, grid, smem, stream)
    if (nsplit == 1) {
        switch (G) {
            case 4: TL_GRP_LAUNCH(1, 4);
            case 3: TL_GRP_LAUNCH(1, 3);
            case 2: TL_GRP_LAUNCH(1, 2);
            default: TL_GRP_LAUNCH(1, 1);
        }
    }
    switch (G) {
        case 4: TL_GRP_LAUNCH(2, 4);
        case 3: TL_GRP_LAUNCH(2, 3);
        case 2: TL_GRP_LAUNCH(2, 2);
        default: TL_GRP_LAUNCH(2, 1);
    }
#undef TL_GRP_LAUNCH
}

}  // namespace tl

// debug: copy the timeline trace of the last TL_GRP_DEBUG=32 launch to the host (roles x TRACE_LEN u64 = tag << 48 | clock)
extern "C" int tl_debug_copy_trace_grp(void* host, size_t bytes) {
    const size_t want = sizeof(unsigned long long) * tl::grp::TRACE_ROLES * tl::grp::TRACE_LEN;
    if (bytes < want) return TL_ERR_ARG;
    cudaDeviceSynchronize();
    return cudaMemcpyFromSymbol(host, tl::grp::g_trace_grp, want) == cudaSuccess ? TL_OK : TL_ERR_CUDA;
}
