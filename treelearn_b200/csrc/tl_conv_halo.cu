// tcgen05 submanifold sparse convolution with a per-tile HALO cache, sm_100a only.  Modes TL_MODE_F16 / TL_MODE_F16X2.
//
// Why (profiles/r02_conv_history.md): every earlier kernel of this library gathers the 128 x 27 neighbour rows of an
// output tile row by row from L2 -- 3456 row requests per tile and 32-channel block, of which 45 % are live and only
// ~230 are DISTINCT (voxels are Morton-sorted, so a tile's 27-neighbourhoods overlap: measured 6.9 live pairs per distinct
// row on the synthetic forest tiles, tools/halo_stats.py).  Those kernels end up bound by the LSU / L2 request rate of
// the gather (~14 B/clk/SM) and by the per-chunk barrier hand-offs around it, not by HBM or the tensor core.
//
// This kernel fetches each distinct row ONCE per tile into shared memory (the tile's halo: tl_halo_build lists the
// distinct rows per tile and rewrites the rulebook as 16-bit indices into that list) and builds the A operand of every
// kernel offset from shared memory:
//   * a GROUP (4 warps = the 4 TMEM lane quarters, + 1 MMA warp) owns one 128-row tile at a time; G groups per CTA, one
//     persistent CTA per SM; warp 5G streams the weights (TMA bulk copies into a ring shared by the groups) or, when the
//     layer's weights fit, they stay resident;
//   * per (segment, 32-channel block) = K-slice: the 4 warps copy the halo rows' 64 B (f16) / 128 B (f16x2) slices with
//     cp.async (one request per distinct row instead of 6.9), then for each of the 27 offsets every lane reads ITS row's
//     slice from shared memory (4 / 8 LDS.128, absent neighbour = predicated off) and writes it to tensor memory with one
//     tcgen05.st.32x32b.x16 per term: lane = row, 16 columns = 32 fp16 K values;
//   * the reads are bank-conflict free by construction ("class swizzle", see tl_halo_build below): rows are stored
//     XOR-swizzled by the voxel's parity class and the tile's rows are permuted over the lanes so that the 8 lanes of an
//     LDS.128 phase hold 8 different classes -- for every one of the 27 offsets at once;
//   * the MMA warp issues TS-form tcgen05.mma (A from tensor memory: 16.6 cycles per N = 32, K = 16 instead of 41 for
//     the shared-memory form, profiles/r02_ts_probe.txt) for a FILL of chunks per barrier round trip (4 chunks in f16, 2
//     in f16x2; two fills in flight), one tcgen05.commit per fill;
//   * epilogue as in tl_conv_grp.cu: tcgen05.ld.16x256b, residual / scale / shift / ReLU, P-layout vector stores (lanes map
//     back to tile rows through the permutation);
//   * levels with fewer tiles than half the SMs split the K-slices of a tile over CTAs (fp32 red.add into a scratch tensor,
//     k_splitk_epilogue_p applies the epilogue).
// The floor of this design is the SM's load/store datapath: every A byte crosses it twice (LDS shared memory -> registers,
// tcgen05.st registers -> tensor memory), 2 x 221 KB per tile and K-slice in f16 (profiles/r02_conv_history.md).
//
// Tensors use the P-layout channel order of treelearn_b200/sparse.py (see tl_conv_grp.cu); weights are the same
// [n_off][c_in/32][(hi, lo)][C_out][32] SWIZZLE_64B images (sparse.pack_weight_grp).
#include <cuda_fp16.h>
#include <stdlib.h>

#include "tl_common.cuh"
#include "tl_tc_ptx.cuh"

namespace tl {
namespace halo {

using namespace tl::tc;

constexpr int MAX_G = 3;
constexpr int MAX_NB = 64;                 // weight ring slots
constexpr int NOFF = 27;
constexpr int LIDX_ROWS = NOFF + 1;        // 27 rulebook rows + the lane -> tile row permutation
constexpr int LIDX_BYTES = LIDX_ROWS * BM * 2;  // the tile's rulebook as 16-bit halo entries (position | swizzle key << 13), per LANE
constexpr uint32_t POS_MASK = 0x1FFFu;
constexpr int ROW_ID_BITS = 28;            // halo_rows entry = row id | class << 28 (-1 = unused position)
constexpr int TMEM_COLS = 512;

struct Launch {
    int num_tiles, rounds;
    int n_slices;        // K-slices = sum over segments of c_in / 32
    int halo_rows;       // rows of a group's halo buffer (zero row + the level's largest halo, rounded up)
    int resident;        // 1: all weight slabs stay in shared memory; 0: ring of `nb` slabs
    int nb;
    uint32_t b_bytes;
    int group_cols;      // TMEM columns per group: accumulator + 2 fills
    int nbuf;            // halo (and rulebook) buffers per group: 2 = the next slice is fetched while this one is used
    // split-K for levels with fewer tiles than SMs (G == 1, rounds == 1, weight ring): CTA b owns tile b / splits and the K-slices
    // [(b % splits) * slices_per_split, ...); accumulators are added into splitk_ws, k_splitk_epilogue_p applies the epilogue
    int splits, slices_per_split;
    float* splitk_ws;
};

struct Layout {
    uint32_t b0, grp0, group_bytes, bars, tmem_slot;
    __device__ __forceinline__ uint32_t a_full(int g, uint32_t s) const { return bars + 8u * (uint32_t)(g * 2 + (int)s); }
    __device__ __forceinline__ uint32_t a_empty(int g, uint32_t s) const { return bars + 8u * (uint32_t)(2 * MAX_G + g * 2 + (int)s); }
    __device__ __forceinline__ uint32_t acc_full(int g) const { return bars + 8u * (uint32_t)(4 * MAX_G + g); }
    __device__ __forceinline__ uint32_t b_full(uint32_t s) const { return bars + 8u * (5 * MAX_G + s); }
    __device__ __forceinline__ uint32_t b_empty(uint32_t s) const { return bars + 8u * (5 * MAX_G + MAX_NB + s); }
    __device__ __forceinline__ uint32_t wres() const { return bars + 8u * (5 * MAX_G + 2 * MAX_NB); }
};
constexpr int BAR_BYTES = (8 * (5 * MAX_G + 2 * MAX_NB + 1) + 15) & ~15;

static inline size_t halo_buf_bytes(int halo_rows, int row_bytes) { return ((size_t)halo_rows * row_bytes + 127) & ~(size_t)127; }
static inline size_t group_bytes(int halo_rows, int row_bytes, int nbuf) {
    return (size_t)nbuf * (halo_buf_bytes(halo_rows, row_bytes) + LIDX_BYTES);
}
// [weights][per group: halo rows, lidx][barriers][tmem slot]
__device__ __forceinline__ Layout carve(uint32_t base, uint32_t b_bytes, int groups, uint32_t gbytes) {
    Layout L;
    L.b0 = base;
    L.grp0 = (L.b0 + b_bytes + 1023u) & ~1023u;
    L.group_bytes = gbytes;
    L.bars = L.grp0 + (uint32_t)groups * gbytes;
    L.tmem_slot = L.bars + BAR_BYTES;
    return L;
}
static inline size_t smem_bytes(size_t b_bytes, int groups, size_t gbytes) {
    return 1024 + ((b_bytes + 1023) & ~(size_t)1023) + (size_t)groups * gbytes + BAR_BYTES + 32;
}

__device__ __forceinline__ void bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ uint32_t ld_shared_u16(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
// 16 B of a halo row, or zeros for an absent neighbour (row == 0): predicated off, the lane takes no bank of the wavefront
__device__ __forceinline__ uint4 ld_rowpiece(uint32_t row, uint32_t off) {
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "@p ld.shared.v4.u32 {%0, %1, %2, %3}, [%5];\n"
        "}\n"
        : "+r"(v.x), "+r"(v.y), "+r"(v.z), "+r"(v.w)
        : "r"(row), "r"(row + off));
    return v;
}
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
// 32 lanes x 32 bit x 16: register i of lane t -> column i of TMEM lane (lane field) + t
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint4& a, const uint4& b, const uint4& c, const uint4& e) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
                 "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w), "r"(c.x), "r"(c.y), "r"(c.z), "r"(c.w),
                 "r"(e.x), "r"(e.y), "r"(e.z), "r"(e.w)
                 : "memory");
}
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

// Warps 0 .. 4G-1: group g = warp / 4 builds the A operand of its tile (TMEM lane quarter = warp % 4) and runs the epilogue.
// Warps 4G .. 5G-1: MMA issue of group warp - 4G.  Warp 5G: TMEM allocation + the weight stream.
// Tile of (round r, CTA c, group g) = (r * gridDim.x + c) * G + g.
// FILL = chunks per barrier round trip (f16: 4 or 2, f16x2: 2 or 1): 16 NSPLIT FILL tensor-memory columns per fill, two fills per group.
template <int NSPLIT, int G, bool RESIDENT, int FILL_>
__global__ void __launch_bounds__(32 * (5 * G + 1), 1) k_conv_halo(const tl_conv_desc d, const Launch P) {
    constexpr uint32_t ROWB = 64u * NSPLIT;                   // bytes of one halo row slice: 32 channels (x hi, lo)
    constexpr uint32_t PIECES = ROWB / 16u;
    constexpr uint32_t FILL = (uint32_t)FILL_;
    constexpr uint32_t A_COLS = 16u * NSPLIT * FILL;          // tensor-memory columns of one fill
    constexpr uint32_t NFILL = (NOFF + FILL - 1) / FILL;
    constexpr uint32_t CHUNK_COLS = 16u * NSPLIT;
    constexpr int kAuxWarp = 5 * G;
    constexpr int FMT = NSPLIT == 2 ? FMT_F16X2 : FMT_F16;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int N = d.c_out;
    const uint32_t NB = (uint32_t)P.nb;
    const uint32_t hbytes = (uint32_t)(((size_t)P.halo_rows * ROWB + 127) & ~(size_t)127);
    const uint32_t gbytes = (uint32_t)P.nbuf * (hbytes + LIDX_BYTES);
    const Layout L = carve(base, P.b_bytes, G, gbytes);
    const uint32_t slab = (uint32_t)N * 64u * NSPLIT;         // weight bytes of one chunk: [C_out][32] fp16 (x hi, lo)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (L.tmem_slot - smem_u32(smem_raw)));
    const uint32_t chunks_total = (uint32_t)P.n_slices * NOFF;

    if (threadIdx.x == 0) {
        for (int g = 0; g < G; ++g) {
            for (uint32_t s = 0; s < 2; ++s) {
                mbar_init(L.a_full(g, s), 4);       // one lane of each of the group's 4 warps: its tcgen05.st have completed
                mbar_init(L.a_empty(g, s), 1);      // tcgen05.commit of the MMAs that read the fill
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

    // K-slice j -> (segment, k-block); TL_MAX_SEG == 3
    uint32_t kb_cnt[TL_MAX_SEG];
#pragma unroll
    for (int s = 0; s < TL_MAX_SEG; ++s) kb_cnt[s] = s < d.n_seg ? (uint32_t)d.seg[s].c_in / 32u : 0u;
    // K-slice range of this CTA and the (segment, k-block, resident slab base) of its first slice
    const bool SK = P.splits > 1;
    const uint32_t j_begin = SK ? (uint32_t)((int)blockIdx.x % P.splits) * (uint32_t)P.slices_per_split : 0u;
    const uint32_t j_end = SK ? min(j_begin + (uint32_t)P.slices_per_split, (uint32_t)P.n_slices) : (uint32_t)P.n_slices;
    uint32_t s_begin = 0, kb_begin = 0, seg_base_begin = 0;
    for (uint32_t jj = 0; jj < j_begin; ++jj)
        if (++kb_begin == kb_cnt[s_begin]) seg_base_begin += (uint32_t)NOFF * kb_cnt[s_begin], kb_begin = 0, ++s_begin;

    if (warp == kAuxWarp) {
        if (RESIDENT) {
            // all slabs, in the order of the global tensors: segment s, offset k, k-block kb -> slab seg_base + k * KB_s + kb
            if (elect_one()) {
                uint32_t total = 0;
                for (int s = 0; s < d.n_seg; ++s) total += (uint32_t)NOFF * kb_cnt[s] * slab;
                mbar_arrive_expect_tx(L.wres(), total);
                uint32_t dst = L.b0;
                for (int s = 0; s < d.n_seg; ++s) {
                    const uint32_t bytes = (uint32_t)NOFF * kb_cnt[s] * slab;
                    const char* src = reinterpret_cast<const char*>(d.seg[s].weight);
                    for (uint32_t off = 0; off < bytes; off += 32768u) bulk_g2s(dst + off, src + off, min(32768u, bytes - off), L.wres());
                    dst += bytes;
                }
            }
            __syncwarp();
        } else {
            // one weight stream for the whole CTA in the order the groups consume it: slice j (segment, k-block), offset k
            const uint32_t total = (uint32_t)P.rounds * (j_end - j_begin) * NOFF;
            uint32_t slot = 0, phase = 0, j = j_begin, k = 0, s = s_begin, kb = kb_begin;
            for (uint32_t p = 0; p < total; ++p) {
                mbar_wait(L.b_empty(slot), phase ^ 1u);
                if (elect_one()) {
                    const char* src = reinterpret_cast<const char*>(d.seg[s].weight) + (size_t)(k * kb_cnt[s] + kb) * slab;
                    mbar_arrive_expect_tx(L.b_full(slot), slab);
                    bulk_g2s(L.b0 + slot * slab, src, slab, L.b_full(slot));
                }
                __syncwarp();
                if (++k == NOFF) {
                    k = 0;
                    if (++kb == kb_cnt[s]) kb = 0, ++s;
                    if (++j == (uint32_t)P.n_slices) j = 0, s = 0, kb = 0;
                }
                if (++slot == NB) slot = 0, phase ^= 1u;
            }
        }
    } else if (warp >= 4 * G) {
        // ===================== MMA issue for group g ===================================================
        const int g = warp - 4 * G;
        const uint32_t acc_col = tmem_base + (uint32_t)(g * P.group_cols);
        const uint32_t a_col0 = acc_col + (uint32_t)N;
        const uint32_t idesc = make_idesc(N, true);
        const uint64_t bdesc0 = make_smem_desc(L.b0, 64);
        const uint32_t slab16 = slab >> 4, half16 = ((uint32_t)N * 64u) >> 4;
        if (RESIDENT) mbar_wait(L.wres(), 0u);
        uint32_t a_slot = 0, a_phase = 0, b_slot = 0, b_phase = 0;
        for (int r = 0; r < P.rounds; ++r) {
            const int tile = SK ? (int)blockIdx.x / P.splits : (r * (int)gridDim.x + (int)blockIdx.x) * G + g;
            const bool valid = tile < P.num_tiles;
            uint32_t seg_base = seg_base_begin, s = s_begin, kb = kb_begin;                // resident slab numbering
            for (uint32_t j = j_begin; j < j_end; ++j) {
                for (uint32_t f = 0; f < NFILL; ++f) {
                    if (valid) {
                        mbar_wait(L.a_full(g, a_slot), a_phase);
                        tc_fence_after();
                    }
#pragma unroll
                    for (uint32_t c = 0; c < FILL; ++c) {
                        const uint32_t k = f * FILL + c;
                        if (k >= (uint32_t)NOFF) break;
                        uint32_t boff;
                        if (RESIDENT) {
                            boff = (seg_base + k * kb_cnt[s] + kb) * slab16;
                        } else {
                            mbar_wait(L.b_full(b_slot), b_phase);        // taken after the A wait: no ring slot is held while waiting
                            boff = b_slot * slab16;
                        }
                        if (elect_one()) {
                            if (valid) {
                                const uint32_t a_col = a_col0 + a_slot * A_COLS + c * CHUNK_COLS;
                                const uint64_t bd = bdesc0 + (uint64_t)boff;
#pragma unroll
                                for (uint32_t kk = 0; kk < 2; ++kk) {
                                    umma_f16_ts(acc_col, a_col + 8 * kk, bd + 2 * kk, idesc, ((j - j_begin) | k | kk) ? 1u : 0u);
                                    if (NSPLIT == 2) {
                                        umma_f16_ts(acc_col, a_col + 8 * kk, bd + half16 + 2 * kk, idesc, 1u);      // hi x lo
                                        umma_f16_ts(acc_col, a_col + 16 + 8 * kk, bd + 2 * kk, idesc, 1u);          // lo x hi
                                    }
                                }
                                if (!RESIDENT) umma_commit(L.b_empty(b_slot));
                                if (c + 1 == FILL || k + 1 == (uint32_t)NOFF) umma_commit(L.a_empty(g, a_slot));
                            } else if (!RESIDENT) {
                                mbar_arrive(L.b_empty(b_slot));           // no tile this round: pass the stream on
                            }
                        }
                        __syncwarp();
                        if (!RESIDENT && ++b_slot == NB) b_slot = 0, b_phase ^= 1u;
                    }
                    if (valid && ++a_slot == 2) a_slot = 0, a_phase ^= 1u;
                }
                if (++kb == kb_cnt[s]) seg_base += (uint32_t)NOFF * kb_cnt[s], kb = 0, ++s;
            }
            if (valid) {
                if (elect_one()) umma_commit(L.acc_full(g));
                __syncwarp();
            }
        }
    } else {
        // ===================== a group's 4 warps: halo -> tensor memory, epilogue =====================
        const int g = warp >> 2, qtr = warp & 3;
        const int q = lane & 3, rr = lane >> 2;          // epilogue: piece q of rows rr + 8 i of my lane quarter
        const int bar_id = 1 + g;
        const int tid = qtr * 32 + lane;                 // 0 .. 127 within the group
        const uint32_t lane_field = ((uint32_t)qtr * 32u) << 16;
        const uint32_t acc_col = tmem_base + (uint32_t)(g * P.group_cols);
        const uint32_t a_col0 = acc_col + (uint32_t)N + lane_field;
        // [halo buffer 0][halo buffer 1]?[rulebook buffer 0][rulebook buffer 1]?
        const uint32_t hbuf0 = L.grp0 + (uint32_t)g * L.group_bytes;
        const uint32_t lbuf0 = hbuf0 + (uint32_t)P.nbuf * hbytes;
        const bool dbl = P.nbuf == 2;
        // row 0 of a halo buffer = zeros (absent neighbours), never overwritten
        if (tid < (int)PIECES) st_shared_v4(hbuf0 + 16u * (uint32_t)tid, 0u, 0u, 0u, 0u);
        if (dbl && tid < (int)PIECES) st_shared_v4(hbuf0 + hbytes + 16u * (uint32_t)tid, 0u, 0u, 0u, 0u);

        // cp.async of the halo slice (segment s, k-block kb) of `tile` (+ its 16-bit rulebook when with_lidx)
        auto load_halo = [&](int tile, int s, uint32_t kb, bool with_lidx, uint32_t hbuf, uint32_t lbuf) {
            const int cnt = __ldg(d.halo_cnt + tile);
            const int32_t* rows = d.halo_rows + (size_t)tile * d.halo_cap;
            const char* src = reinterpret_cast<const char*>(d.seg[s].src) + kb * ROWB;
            const size_t rb = (size_t)d.seg[s].src_stride * (2u * NSPLIT);
            // row ids first, eight independent loads at a time (a load -> copy chain per row costs a global-memory latency
            // per iteration: 24 % of all stall samples in profiles/r02_ncu_halo_v1_c32.txt), then the copies
            constexpr int RPI = 128 / (int)PIECES;          // rows per pass of the group's 128 threads
            const int rsub = tid / (int)PIECES;
            const uint32_t p = (uint32_t)tid % PIECES;
            for (int r0 = 0; r0 < cnt; r0 += 8 * RPI) {
                int rid[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int row = r0 + u * RPI + rsub;
                    rid[u] = row < cnt ? __ldg(rows + row) : -1;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (rid[u] < 0) continue;
                    const uint32_t hr = (uint32_t)(r0 + u * RPI + rsub) + 1u;
                    const uint32_t cls = (uint32_t)rid[u] >> ROW_ID_BITS;                  // the row's parity class = its swizzle key
                    const uint32_t sw = NSPLIT == 1 ? (cls & 3u) : cls;                    // (f16: bit 2 of the class is the position's parity)
                    cp_async16_cg(hbuf + hr * ROWB + ((p ^ sw) << 4), src + (size_t)(rid[u] & ((1 << ROW_ID_BITS) - 1)) * rb + p * 16u, 16u);
                }
            }
            if (with_lidx) {
                const char* lsrc = reinterpret_cast<const char*>(d.halo_lidx) + (size_t)tile * LIDX_BYTES;
                for (int i = tid; i < LIDX_BYTES / 16; i += 128) cp_async16_cg(lbuf + 16u * (uint32_t)i, lsrc + 16 * i, 16u);
            }
            cp_async_commit();
        };

        uint32_t a_slot = 0, a_phase = 0;
        uint32_t hsel = 0, lsel = 0;                     // buffer of the current slice / of the current tile's rulebook
        {
            const int tile0 = SK ? (int)blockIdx.x / P.splits : (int)blockIdx.x * G + g;
            if (tile0 < P.num_tiles) load_halo(tile0, (int)s_begin, kb_begin, true, hbuf0, lbuf0);
        }
        for (int r = 0; r < P.rounds; ++r) {
            const int tile = SK ? (int)blockIdx.x / P.splits : (r * (int)gridDim.x + (int)blockIdx.x) * G + g;
            if (tile >= P.num_tiles) continue;           // (the group's MMA warp passes the weight stream on)
            const int next_tile = SK ? P.num_tiles : ((r + 1) * (int)gridDim.x + (int)blockIdx.x) * G + g;
            int s = (int)s_begin;
            uint32_t kb = kb_begin;
            int prow[4] = {0, 0, 0, 0};
            const uint32_t lbuf = lbuf0 + lsel * LIDX_BYTES;
            for (uint32_t j = j_begin; j < j_end; ++j) {
                const uint32_t hbuf = hbuf0 + hsel * hbytes;
                // the slice after this one: the tile's next K-slice, or the next tile's first one (with its rulebook)
                const bool last = j + 1 == j_end;
                int ns = s;
                uint32_t nkb = kb + 1;
                if (nkb == kb_cnt[s]) nkb = 0, ++ns;
                auto prefetch = [&]() {
                    const uint32_t nh = hbuf0 + (dbl ? (hsel ^ 1u) : 0u) * hbytes;
                    if (!last) load_halo(tile, ns, nkb, false, nh, 0u);
                    else if (next_tile < P.num_tiles) load_halo(next_tile, 0, 0u, true, nh, lbuf0 + (dbl ? (lsel ^ 1u) : 0u) * LIDX_BYTES);
                };
                cp_async_wait_all();
                bar_sync(bar_id, 128);       // this slice (and the rulebook) has landed for all 4 warps; the previous slice is read out
                if (j == j_begin) {          // tile rows of the 4 TMEM lanes this thread writes out (the rulebook buffer may be refilled before the epilogue)
#pragma unroll
                    for (int i = 0; i < 4; ++i) prow[i] = (int)ld_shared_u16(lbuf + 2u * (uint32_t)(NOFF * BM + qtr * 32 + rr + 8 * i));
                }
                if (dbl) prefetch();         // into the other buffer, while this slice is used
                const uint32_t lrow = lbuf + 2u * (uint32_t)tid;
#pragma unroll 1
                for (uint32_t f = 0; f < NFILL; ++f) {
                    // the shared-memory reads do not depend on the fill's tensor-memory slot: issue the first ones before the wait
                    // (measured, round 2: fetching the rulebook entries one fill ahead and the halo row ids two tiles ahead
                    // into shared memory changes nothing -- 14.98 vs 14.97 ms of conv time per step -- although 20 % of the
                    // stall samples sit on those loads: the groups hide each other's latencies)
                    uint32_t rowa[FILL], swz[FILL];
#pragma unroll
                    for (uint32_t c = 0; c < FILL; ++c) {
                        const uint32_t k = min(f * FILL + c, (uint32_t)NOFF - 1u);
                        const uint32_t li = ld_shared_u16(lrow + k * (2u * BM));
                        rowa[c] = li ? hbuf + (li & POS_MASK) * ROWB : 0u;          // 0 = absent neighbour: zeros without touching shared memory
                        swz[c] = NSPLIT == 1 ? ((li >> 13) & 3u) : (li >> 13);
                    }
                    const uint32_t ta = a_col0 + a_slot * A_COLS;
                    if (NSPLIT == 1) {           // two chunks (8 LDS.128) in flight
                        uint4 v[2][4];
#pragma unroll
                        for (uint32_t c = 0; c < 2; ++c)
#pragma unroll
                            for (uint32_t i = 0; i < 4; ++i) v[c][i] = ld_rowpiece(rowa[c], (i ^ swz[c]) << 4);
                        mbar_wait(L.a_empty(g, a_slot), a_phase ^ 1u);
                        tc_fence_after();
                        tmem_st_32x32b_x16(ta, v[0][0], v[0][1], v[0][2], v[0][3]);
                        if (f * FILL + 1 < (uint32_t)NOFF) tmem_st_32x32b_x16(ta + CHUNK_COLS, v[1][0], v[1][1], v[1][2], v[1][3]);
                        if (FILL == 4) {
#pragma unroll
                            for (uint32_t c = 0; c < 2; ++c)
#pragma unroll
                                for (uint32_t i = 0; i < 4; ++i) v[c][i] = ld_rowpiece(rowa[(FILL > 2 ? 2 : 0) + c], (i ^ swz[(FILL > 2 ? 2 : 0) + c]) << 4);
                            if (f * FILL + 2 < (uint32_t)NOFF) tmem_st_32x32b_x16(ta + 2 * CHUNK_COLS, v[0][0], v[0][1], v[0][2], v[0][3]);
                            if (f * FILL + 3 < (uint32_t)NOFF) tmem_st_32x32b_x16(ta + 3 * CHUNK_COLS, v[1][0], v[1][1], v[1][2], v[1][3]);
                        }
                    } else {                     // one chunk (hi + lo: 8 LDS.128) in flight
                        uint4 v[8];
#pragma unroll
                        for (uint32_t i = 0; i < 8; ++i) v[i] = ld_rowpiece(rowa[0], (i ^ swz[0]) << 4);
                        mbar_wait(L.a_empty(g, a_slot), a_phase ^ 1u);
                        tc_fence_after();
                        tmem_st_32x32b_x16(ta, v[0], v[1], v[2], v[3]);                        // hi
                        tmem_st_32x32b_x16(ta + 16u, v[4], v[5], v[6], v[7]);                  // lo
                        if (FILL == 2 && f * FILL + 1 < (uint32_t)NOFF) {
#pragma unroll
                            for (uint32_t i = 0; i < 8; ++i) v[i] = ld_rowpiece(rowa[FILL - 1], (i ^ swz[FILL - 1]) << 4);
                            tmem_st_32x32b_x16(ta + CHUNK_COLS, v[0], v[1], v[2], v[3]);
                            tmem_st_32x32b_x16(ta + CHUNK_COLS + 16u, v[4], v[5], v[6], v[7]);
                        }
                    }
                    tmem_wait_st();
                    tc_fence_before();
                    if (lane == 0) mbar_arrive(L.a_full(g, a_slot));
                    if (++a_slot == 2) a_slot = 0, a_phase ^= 1u;
                }
                if (!dbl) {
                    bar_sync(bar_id, 128);   // every warp is done reading this slice: the one halo buffer is free
                    prefetch();              // (the next tile's first slice overlaps the accumulator wait + epilogue)
                }
                kb = nkb, s = ns;
                hsel ^= dbl ? 1u : 0u;
            }
            lsel ^= dbl ? 1u : 0u;

            // ---- epilogue: this warp's 32 rows; lane (rr, q) holds positions 8q .. 8q+7 of every 32-channel block of rows
            //      rr, rr + 8 (half 0) and rr + 16, rr + 24 (half 1) of its quarter
            const int64_t trow0 = (int64_t)tile * BM;       // TMEM lane qtr * 32 + rr + 8 i holds tile row prow[i]
            const bool has_res = d.residual != nullptr && !SK;      // split-K: the second pass adds the residual
            float4 res[4][2];
            auto fetch_residual = [&](int c0) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int64_t grow = trow0 + prow[i];
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
            for (int c0 = 0; c0 < N; c0 += 32) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {          // lanes 16h .. 16h+15 of the quarter: rows rr + 16h and rr + 16h + 8
                    uint32_t v[16];
                    tmem_ld_16x256b_x4(acc_col + lane_field + ((uint32_t)(16 * h) << 16) + (uint32_t)c0, v);
                    tmem_wait_ld();
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
                    if (SK) {                // partial sums of this CTA's K-slices -> fp32 scratch (P-layout positions, like out_raw)
#pragma unroll
                        for (int ab = 0; ab < 2; ++ab) {
                            const int64_t grow = trow0 + prow[2 * h + ab];
                            if (grow >= d.n_out) continue;
                            float* p = P.splitk_ws + grow * N + c0 + 8 * q;
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(x[ab][0]), "f"(x[ab][1]), "f"(x[ab][2]), "f"(x[ab][3]) : "memory");
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p + 4), "f"(x[ab][4]), "f"(x[ab][5]), "f"(x[ab][6]), "f"(x[ab][7]) : "memory");
                        }
                        continue;
                    }
                    if (d.out_raw) {
#pragma unroll
                        for (int ab = 0; ab < 2; ++ab) {
                            const int64_t grow = trow0 + prow[2 * h + ab];
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
                            const int64_t grow = trow0 + prow[2 * h + ab];
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
            tc_fence_before();      // accumulator reads are complete (wait::ld) before this warp's next a_full arrival
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kAuxWarp) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// split-K second pass over P-layout tensors: v = scratch (+ residual) -> raw / activated outputs; position m of a
// 32-channel block holds logical channel p_chan(m) (scale / shift vectors are in logical order)
template <int FMT>
__global__ void __launch_bounds__(256) k_splitk_epilogue_p(const tl_conv_desc d, const float* __restrict__ ws) {
    const int64_t total = (int64_t)d.n_out * d.c_out;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int col = (int)(e % d.c_out);
        const int ch = (col & ~31) + p_chan(col & 31);
        float v = ws[e];
        if (d.residual) v += __ldg(d.residual + e);
        if (d.out_raw) d.out_raw[e] = v;
        if (d.out_act1) store_act1<FMT>(d.out_act1, e, d.c_out, fmaxf(fmaf(v, __ldg(d.scale1 + ch), __ldg(d.shift1 + ch)), 0.f));
        if (d.out_act2) store_act1<FMT>(d.out_act2, e, d.c_out, fmaxf(fmaf(v, __ldg(d.scale2 + ch), __ldg(d.shift2 + ch)), 0.f));
    }
}

template <int NSPLIT, int G, bool RESIDENT, int FILL>
static int launch(const tl_conv_desc& d, const Launch& P, int grid, size_t smem, cudaStream_t stream) {
    static bool configured[16] = {false};
    int dev = 0;
    TL_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 16 && !configured[dev]) {     // function attributes are per device
        TL_CUDA_CHECK(cudaFuncSetAttribute(k_conv_halo<NSPLIT, G, RESIDENT, FILL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured[dev] = true;
    }
    k_conv_halo<NSPLIT, G, RESIDENT, FILL><<<grid, 32 * (5 * G + 1), smem, stream>>>(d, P);
    TL_LAUNCH_CHECK();
    return TL_OK;
}

// ---- tl_halo_build: per 128-row tile, the list of distinct neighbour rows and the rulebook as entries into it ----------
// Bank-conflict-free by construction (profiles/r02_conv_history.md, "class swizzle"): a voxel's PARITY CLASS
// cls = (x & 1) << 2 | (y & 1) << 1 | (z & 1) is the low 3 bits of its Morton key.  A halo row is stored with its 16 B pieces
// XOR-swizzled by its class (the position's parity supplies bit 2 for 64 B rows), and the tile's rows are PERMUTED over the
// 128 lanes so that the 8 lanes of a shared-memory phase hold 8 different classes wherever the tile's class histogram
// allows (<= 16 rows per class).  The neighbours at one kernel offset of 8 voxels of different classes again have 8
// different classes (a translation flips the same parity bits for all of them), so the 8 lanes of every LDS.128 phase of
// the conv kernel hit 8 different bank groups for every one of the 27 offsets.
//   halo_rows[tile][p - 1] = row id | cls << 28 at position p >= 1 (-1 = unused position; p is odd iff cls bit 2 is set)
//   lidx[tile][k][lane]    = p | cls << 13 of the neighbour at offset k of the tile row held by `lane` (0 = absent)
//   lidx[tile][27][lane]   = that tile row (lane -> row permutation)
constexpr int HB_THREADS = 256;
constexpr int HB_TABLE = 4096;       // open-addressing table, > the 3456 entries of a tile
constexpr int HB_SORT = 2048;        // largest `cap`
constexpr int HB_PER = (NOFF * BM + HB_THREADS - 1) / HB_THREADS;      // rulebook entries per thread (14)

__global__ void __launch_bounds__(HB_THREADS) k_halo_build(const int32_t* __restrict__ nbr, const uint64_t* __restrict__ keys, int64_t n,
                                                            int64_t stride, int cap, int32_t* __restrict__ halo_rows,
                                                            int32_t* __restrict__ halo_cnt, uint16_t* __restrict__ lidx,
                                                            int32_t* __restrict__ max_cnt) {
    __shared__ int table[HB_TABLE];
    __shared__ unsigned short rank[HB_TABLE];      // table slot -> position | class << 13 of its row
    __shared__ int list[HB_SORT];
    __shared__ unsigned short pos_list[HB_SORT];   // sorted-list index -> position | class << 13
    __shared__ int warp_tot[HB_THREADS / 32];
    __shared__ int total_s, used_s;
    __shared__ unsigned char cls_row[BM], lane_of[BM], taken[BM], hole_list[BM], cls_list[HB_SORT];
    __shared__ unsigned char cnt_wc[BM / 32][8], left_w[BM / 32], hole_w[BM / 32];
    const int tile = blockIdx.x, t = threadIdx.x;
    const int64_t row0 = (int64_t)tile * BM;
    for (int i = t; i < HB_TABLE; i += HB_THREADS) table[i] = -1;
    if (t < BM) {
        cls_row[t] = row0 + t < n ? (unsigned char)(__ldg(keys + row0 + t) & 7u) : (unsigned char)(t & 7);
        taken[t] = 0;
        if (t < (BM / 32) * 8) (&cnt_wc[0][0])[t] = 0;
    }
    // this thread's entries e = t + 256 i (offset e / 128, row e % 128): loaded once, kept in registers with their table slot
    int val[HB_PER];
    unsigned short slot[HB_PER];
#pragma unroll
    for (int i = 0; i < HB_PER; ++i) {
        const int e = t + HB_THREADS * i;
        const int k = e / BM, r = e % BM;
        val[i] = (e < NOFF * BM && row0 + r < n) ? __ldg(nbr + (size_t)k * stride + row0 + r) : -1;
    }
    __syncthreads();
    // ---- lane permutation: the j-th row of class c takes lane 8 j + c (phase j); rows beyond 16 of a class fill the holes
    const uint32_t lt = (1u << (t & 31)) - 1u;
    const int my_cls = t < BM ? cls_row[t] : 0;
    int rk = 0;
    if (t < BM) {       // warps 0-3, whole warps
        const uint32_t same = __match_any_sync(0xffffffffu, my_cls);
        rk = __popc(same & lt);
        if (rk == 0) cnt_wc[t >> 5][my_cls] = (unsigned char)__popc(same);
    }
    __syncthreads();
    if (t < BM) {
        for (int w = 0; w < (t >> 5); ++w) rk += cnt_wc[w][my_cls];
        if (rk < 16) {
            lane_of[t] = (unsigned char)(8 * rk + my_cls);
            taken[8 * rk + my_cls] = 1;
        }
    }
#pragma unroll
    for (int i = 0; i < HB_PER; ++i) {
        const int v = val[i];
        if (v < 0) continue;
        uint32_t h = ((uint32_t)v * 2654435761u) >> 20;          // 12 bits
        while (true) {
            const int old = atomicCAS(&table[h], -1, v);
            if (old == -1 || old == v) break;
            h = (h + 1) & (HB_TABLE - 1);
        }
        slot[i] = (unsigned short)h;
    }
    __syncthreads();
    // the i-th left-over row (row order) takes the i-th free lane (lane order)
    const bool is_left = t < BM && rk >= 16, is_hole = t < BM && !taken[t];
    uint32_t bl = 0, bh = 0;
    if (t < BM) {
        bl = __ballot_sync(0xffffffffu, is_left), bh = __ballot_sync(0xffffffffu, is_hole);
        if ((t & 31) == 0) left_w[t >> 5] = (unsigned char)__popc(bl), hole_w[t >> 5] = (unsigned char)__popc(bh);
    }
    __syncthreads();
    int pl = __popc(bl & lt), ph = __popc(bh & lt);
    if (t < BM) {
        for (int w = 0; w < (t >> 5); ++w) pl += left_w[w], ph += hole_w[w];
        if (is_hole) hole_list[ph] = (unsigned char)t;
    }
    __syncthreads();
    if (is_left) lane_of[t] = hole_list[pl];
    // compact the occupied slots: each thread owns the HB_TABLE / HB_THREADS slots t, t + 256, ...
    constexpr int PER = HB_TABLE / HB_THREADS;
    int mine = 0;
#pragma unroll
    for (int i = 0; i < PER; ++i) mine += table[i * HB_THREADS + t] >= 0;      // interleaved slots: conflict-free (t * PER + i was 16-way)
    int incl = mine;
    for (int off = 1; off < 32; off <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, off);
        if ((t & 31) >= off) incl += v;
    }
    if ((t & 31) == 31) warp_tot[t >> 5] = incl;
    __syncthreads();
    if (t == 0) {
        int acc = 0;
        for (int w = 0; w < HB_THREADS / 32; ++w) {
            const int v = warp_tot[w];
            warp_tot[w] = acc;
            acc += v;
        }
        total_s = acc;
    }
    __syncthreads();
    const int total = total_s;
    if (total > cap) {                           // the caller sees max_cnt > cap and does not use the halo kernel
        if (t == 0) halo_cnt[tile] = 0, atomicMax(max_cnt, total);
        return;
    }
    int at = warp_tot[t >> 5] + incl - mine;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int v = table[i * HB_THREADS + t];
        if (v >= 0) list[at++] = v;
    }
    __syncthreads();
    // (no sort: list order = hash-slot order.  Positions are free since the class swizzle; ascending row ids only made the
    // halo fetch walk DRAM in order, and the ~230 rows of a tile sit in a few KB-sized runs either way)
    // ---- positions: rows whose class has bit 2 clear take the even positions 2, 4, ..., the others the odd ones 1, 3, ...
    for (int i = t; i < total; i += HB_THREADS) cls_list[i] = (unsigned char)(__ldg(keys + list[i]) & 7u);
    __syncthreads();
    if (t < 32) {
        int n0 = 0, n1 = 0;
        for (int base = 0; base < total; base += 32) {
            const int i = base + t;
            const uint32_t cv = i < total ? (uint32_t)cls_list[i] : 0u;
            const bool odd = i < total && (cv & 4u);
            const bool even = i < total && !(cv & 4u);
            const uint32_t bo = __ballot_sync(0xffffffffu, odd), be = __ballot_sync(0xffffffffu, even);
            if (odd) pos_list[i] = (unsigned short)((1 + 2 * (n1 + __popc(bo & lt))) | (cv << 13));
            if (even) pos_list[i] = (unsigned short)((2 + 2 * (n0 + __popc(be & lt))) | (cv << 13));
            n1 += __popc(bo), n0 += __popc(be);
        }
        if (t == 0) used_s = max(2 * n0, 2 * n1 - 1);
    }
    __syncthreads();
    const int used = used_s;
    if (t == 0) {
        halo_cnt[tile] = used <= cap ? used : 0;
        atomicMax(max_cnt, used);
    }
    if (used > cap) return;
    for (int i = t; i < used; i += HB_THREADS) halo_rows[(size_t)tile * cap + i] = -1;
    __syncthreads();
    for (int i = t; i < total; i += HB_THREADS) {
        const int v = list[i];
        const uint32_t pc = pos_list[i];
        halo_rows[(size_t)tile * cap + (pc & POS_MASK) - 1] = v | (int)((pc >> 13) << ROW_ID_BITS);
        uint32_t h = ((uint32_t)v * 2654435761u) >> 20;
        while (table[h] != v) h = (h + 1) & (HB_TABLE - 1);
        rank[h] = (unsigned short)pc;
    }
    __syncthreads();
    uint16_t* out = lidx + (size_t)tile * (LIDX_ROWS * BM);
#pragma unroll
    for (int i = 0; i < HB_PER; ++i) {
        const int e = t + HB_THREADS * i;
        if (e < NOFF * BM) out[(e / BM) * BM + lane_of[e % BM]] = val[i] >= 0 ? rank[slot[i]] : (unsigned short)0;
    }
    if (t < BM) out[NOFF * BM + lane_of[t]] = (unsigned short)t;
}

}  // namespace halo

static int env_int_halo(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

bool conv_halo_eligible(const tl_conv_desc& d, uint32_t src_fp32_mask) {
    if (!d.halo_rows || !d.halo_cnt || !d.halo_lidx || d.halo_cap <= 0 || d.halo_umax <= 0 || d.halo_umax > d.halo_cap) return false;
    if (d.c_out % 32 != 0 || d.c_out > 256 || src_fp32_mask || d.n_seg < 1) return false;
    for (int s = 0; s < d.n_seg; ++s) {
        const tl_conv_seg& g = d.seg[s];
        if (g.n_off != halo::NOFF || !g.index || g.index != d.seg[0].index) return false;
        if (g.c_in % 32 != 0 || g.c_in > 256 || g.src_stride % 8 != 0) return false;
    }
    return true;
}

// nsplit: 1 = TL_MODE_F16, 2 = TL_MODE_F16X2.  Returns TL_ERR_UNSUPPORTED when no shared-memory configuration exists.
int conv_fwd_halo(const tl_conv_desc& d, cudaStream_t stream, int nsplit) {
    static int num_sms[16] = {0};
    int dev = 0;
    TL_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 16) return TL_ERR_UNSUPPORTED;
    if (!num_sms[dev]) TL_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms[dev], cudaDevAttrMultiProcessorCount, dev));
    const int sms = num_sms[dev];
    const int n = d.c_out;
    halo::Launch P;
    memset(&P, 0, sizeof(P));
    P.num_tiles = (d.n_out + tc::BM - 1) / tc::BM;
    for (int s = 0; s < d.n_seg; ++s) P.n_slices += d.seg[s].c_in / 32;
    P.halo_rows = ((d.halo_umax + 1 + 7) / 8) * 8;
    // chunks per fill: the larger fill (fewer barrier round trips) unless the smaller one lets another group fit tensor memory
    const int fill_hi = nsplit == 1 ? 4 : 2, fill_lo = fill_hi / 2;
    auto groups_for = [&](int fill) { const int c = n + 2 * 16 * nsplit * fill; const int g = halo::TMEM_COLS / c; return g > halo::MAX_G ? halo::MAX_G : g; };
    int fill = groups_for(fill_lo) > groups_for(fill_hi) ? fill_lo : fill_hi;
    if (env_int_halo("TL_HALO_FILL", 0) == fill_lo || env_int_halo("TL_HALO_FILL", 0) == fill_hi) fill = env_int_halo("TL_HALO_FILL", 0);
    P.group_cols = n + 2 * 16 * nsplit * fill;
    const int row_bytes = 64 * nsplit;
    const uint32_t slab = (uint32_t)n * 64u * nsplit;
    const size_t gbytes1 = halo::group_bytes(P.halo_rows, row_bytes, 1), gbytes2 = halo::group_bytes(P.halo_rows, row_bytes, 2);
    const size_t budget = (size_t)env_int_halo("TL_HALO_SMEM_KB", 224) * 1024;
    const size_t wbytes = (size_t)P.n_slices * halo::NOFF * slab;
    int G = halo::TMEM_COLS / P.group_cols;
    if (G > halo::MAX_G) G = halo::MAX_G;
    if (env_int_halo("TL_HALO_GROUPS", 0) > 0 && env_int_halo("TL_HALO_GROUPS", 0) < G) G = env_int_halo("TL_HALO_GROUPS", 0);
    while (G > 1 && (G - 1) * sms >= P.num_tiles) --G;          // small levels: spread the tiles over the SMs first
    // levels with fewer tiles than half the SMs (the two or three deepest): one tile is a serial chain of n_slices x 27 chunks
    // of N-wide MMAs on ONE SM (measured 70-120 us per launch for a few thousand voxels) -> split the K-slices over CTAs
    P.splits = 1, P.slices_per_split = P.n_slices;
    if (d.splitk_ws && P.n_slices >= 2 && env_int_halo("TL_HALO_SPLITK", 1) != 0) {
        int sp = sms / P.num_tiles;
        if (sp > P.n_slices) sp = P.n_slices;
        if (sp >= 2) {
            P.slices_per_split = (P.n_slices + sp - 1) / sp;
            P.splits = (P.n_slices + P.slices_per_split - 1) / P.slices_per_split;
        }
    }
    if (P.splits > 1) G = 1;
    // per G: resident weights + two halo buffers, else resident + one, else a weight ring (>= 4 slabs) + two, else ring + one
    int resident = 0, nbuf = 1;
    const bool allow_res = env_int_halo("TL_HALO_RESIDENT", 1) != 0 && P.splits == 1, allow_dbl = env_int_halo("TL_HALO_DOUBLE", 1) != 0;
    for (; G >= 1; --G) {
        if (allow_res && allow_dbl && halo::smem_bytes(wbytes, G, gbytes2) <= budget) { resident = 1, nbuf = 2; break; }
        if (allow_res && halo::smem_bytes(wbytes, G, gbytes1) <= budget) { resident = 1, nbuf = 1; break; }
        if (allow_dbl && halo::smem_bytes((size_t)4 * slab, G, gbytes2) <= budget) { resident = 0, nbuf = 2; break; }
        if (halo::smem_bytes((size_t)4 * slab, G, gbytes1) <= budget) { resident = 0, nbuf = 1; break; }
    }
    if (G < 1) return TL_ERR_UNSUPPORTED;
    const size_t gbytes = nbuf == 2 ? gbytes2 : gbytes1;
    P.nbuf = nbuf;
    P.resident = resident;
    if (resident) {
        P.nb = 0;
        P.b_bytes = (uint32_t)wbytes;
    } else {
        int nb = (int)((budget - halo::smem_bytes(0, G, gbytes)) / slab);
        if (nb > halo::MAX_NB) nb = halo::MAX_NB;
        P.nb = nb;
        P.b_bytes = (uint32_t)nb * slab;
    }
    const size_t smem = halo::smem_bytes(P.b_bytes, G, gbytes);
    int grid = (P.num_tiles + G - 1) / G;
    if (grid > sms) grid = sms;
    P.rounds = (P.num_tiles + grid * G - 1) / (grid * G);
    if (P.splits > 1) {
        grid = P.num_tiles * P.splits, P.rounds = 1, P.splitk_ws = d.splitk_ws;
        TL_CUDA_CHECK(cudaMemsetAsync(d.splitk_ws, 0, sizeof(float) * (size_t)d.n_out * d.c_out, stream));
    }
    auto run = [&]() -> int {
#define TL_HALO_LAUNCH2(NS, GG, FF) return P.resident ? halo::launch<NS, GG, true, FF>(d, P, grid, smem, stream) : halo::launch<NS, GG, false, FF>(d, P, grid, smem, stream)
#define TL_HALO_LAUNCH(NS, GG, FHI, FLO) if (fill == FHI) { TL_HALO_LAUNCH2(NS, GG, FHI); } else { TL_HALO_LAUNCH2(NS, GG, FLO); }
    if (nsplit == 1) {
        switch (G) {
            case 3: TL_HALO_LAUNCH(1, 3, 4, 2);
            case 2: TL_HALO_LAUNCH(1, 2, 4, 2);
            default: TL_HALO_LAUNCH(1, 1, 4, 2);
        }
    }
    switch (G) {
        case 3: TL_HALO_LAUNCH(2, 3, 2, 1);
        case 2: TL_HALO_LAUNCH(2, 2, 2, 1);
        default: TL_HALO_LAUNCH(2, 1, 2, 1);
    }
#undef TL_HALO_LAUNCH2
#undef TL_HALO_LAUNCH
    };
    const int rc = run();
    if (rc != TL_OK || P.splits == 1) return rc;
    const int64_t total = (int64_t)d.n_out * d.c_out;
    const unsigned eg = (unsigned)((total + 255) / 256 > sms * 8 ? sms * 8 : (total + 255) / 256);
    if (nsplit == 1) halo::k_splitk_epilogue_p<tc::FMT_F16><<<eg, 256, 0, stream>>>(d, d.splitk_ws);
    else halo::k_splitk_epilogue_p<tc::FMT_F16X2><<<eg, 256, 0, stream>>>(d, d.splitk_ws);
    TL_LAUNCH_CHECK();
    return TL_OK;
}

}  // namespace tl

extern "C" int tl_halo_build(const int32_t* nbr, const uint64_t* keys, int64_t n, int64_t nbr_stride, int32_t cap, int32_t* halo_rows,
                             int32_t* halo_cnt, uint16_t* halo_lidx, int32_t* max_cnt, void* stream) {
    if (n <= 0) return TL_OK;
    if (!nbr || !keys || n >= (1ll << tl::halo::ROW_ID_BITS) || !halo_rows || !halo_cnt || !halo_lidx || !max_cnt || cap <= 0 || cap > tl::halo::HB_SORT) {
        tl::set_error("tl_halo_build: bad argument");
        return TL_ERR_ARG;
    }
    const int tiles = (int)((n + tl::tc::BM - 1) / tl::tc::BM);
    tl::halo::k_halo_build<<<tiles, tl::halo::HB_THREADS, 0, (cudaStream_t)stream>>>(nbr, keys, n, nbr_stride, cap, halo_rows, halo_cnt, halo_lidx, max_cnt);
    TL_LAUNCH_CHECK();
    return TL_OK;
}
