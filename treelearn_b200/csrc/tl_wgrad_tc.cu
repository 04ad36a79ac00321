// tcgen05 weight gradient of the sparse convolutions (training path), sm_100a only.
//
//   dW[k][ci][co] = sum_r src[index[k][r]][ci] * d_out[r][co]          (spconv's wgrad; tools/training/train.py:40 backward)
//
// is, per kernel offset k, a (C_in x rows) . (rows x C_out) product whose REDUCTION dimension is the voxel row.  Both
// operands are stored row-major by voxel, i.e. with the non-reduction dimension contiguous: "MN-major" in UMMA terms,
// which tcgen05.mma takes for TF32 through the a_major / b_major bits of the instruction descriptor and MN-major
// shared-memory descriptors.  So the fp32 tensors feed the tensor core as they lie in HBM -- no transposed copies:
//   * a UNIT = (offset k, 32-channel block kb of C_in); 4 units stack along M = 128 (TMEM lane = 32 unit + ci), a GROUP;
//     a group's accumulator is 128 lanes x N = C_out fp32 columns of tensor memory, floor(512 / N) groups are resident at once
//     (= one PASS over the rows; C = 32: all 27 offsets in one pass, 224 columns);
//   * A stage = the K block of 32 rows for one group: per unit 32 gathered rows x 128 B, written by cp.async straight into
//     the canonical MN-major SWIZZLE_128B_BASE32B image (atoms of 4 rows x 128 B, 32 B chunk j of row r at j ^ (r & 3): the
//     one swizzled layout the tensor core accepts for MN-major 32-bit operands);
//     absent neighbours are zero-filled by cp.async (src-size 0, no global read);
//   * B stage = the same 32 rows of d_out, one 8-row x 128 B atom column per 32 output channels, shared by every group;
//   * one MMA warp issues 4 x tcgen05.mma.kind::tf32 (M128, N = C_out, K8) per A stage; accumulators stay in tensor memory
//     over the CTA's whole row range and are written out ONCE (per-CTA partial sums, then a fixed-order reduction kernel:
//     run-to-run deterministic, unlike the atomicAdd kernel this replaces).
// grid = (row splits, passes); one CTA per SM (512 TMEM columns).  TF32 operands are the upper 19 bits of the fp32 values
// (hardware truncation), fp32 accumulate.
#include <stdlib.h>

#include "tl_common.cuh"
#include "tl_tc_ptx.cuh"

namespace tl {
namespace wg {

using namespace tl::tc;

constexpr int KR = 32;                      // rows per stage (K block)
constexpr uint32_t UNIT_BYTES = KR * 128;   // one unit (or one 32-column block of d_out) of a stage
constexpr uint32_t A_STAGE = 4 * UNIT_BYTES;
constexpr int NA = 6, NBS = 3;
constexpr int TMEM_COLS = 512;
constexpr int THREADS = 160;                // warps 0-3 gather + epilogue, warp 4 MMA issue + TMEM allocation

struct Params {
    const float* src;
    int64_t src_stride;
    int c_in, n_off;
    const int32_t* index;
    int64_t index_stride;
    const uint32_t* tile_mask;
    const float* d_out;
    int64_t n_out;
    int c_out;
    float* out;              // [splits][n_off * c_in * c_out] partial sums (dW itself when splits == 1)
    int units, groups, gp, tiles, tiles_per_split;
    uint32_t lbo_a, lbo_b, sbo, idesc_xor;   // descriptor fields (TL_WG_* debug overrides)
};

struct Layout {
    uint32_t a0, b0, b_bytes, idx0, bars, tmem_slot, flags;
    __device__ __forceinline__ uint32_t a(uint32_t s) const { return a0 + s * A_STAGE; }
    __device__ __forceinline__ uint32_t b(uint32_t s) const { return b0 + s * b_bytes; }
    __device__ __forceinline__ uint32_t a_full(uint32_t s) const { return bars + 8u * s; }
    __device__ __forceinline__ uint32_t a_empty(uint32_t s) const { return bars + 8u * (NA + s); }
    __device__ __forceinline__ uint32_t b_full(uint32_t s) const { return bars + 8u * (2 * NA + s); }
    __device__ __forceinline__ uint32_t b_empty(uint32_t s) const { return bars + 8u * (2 * NA + NBS + s); }
    __device__ __forceinline__ uint32_t acc_full() const { return bars + 8u * (2 * NA + 2 * NBS); }
};
constexpr int BAR_BYTES = 8 * (2 * NA + 2 * NBS + 1) + 8;

// rulebook rows staged per producer warp: [2 buffers][groups per pass][128 rows] int32 (the tile after the current one is in flight)
__host__ __device__ static inline uint32_t idx_bytes(int gp) { return 4u * 2u * (uint32_t)gp * BM * 4u; }
static inline size_t smem_bytes(int c_out, int gp) {
    return 1024 + (size_t)NA * A_STAGE + (size_t)NBS * (c_out / 32) * UNIT_BYTES + idx_bytes(gp) + BAR_BYTES + 32;
}

// MN-major shared-memory matrix descriptor for 32-bit operands.  TF32 MN-major operands have ONE legal swizzled layout,
// SWIZZLE_128B_BASE32B (layout type 1; cutlass sm100_common.inl: "for mn-major tf32 operands, SW128_32B is the only
// available smem layout"): atoms of 4 K rows x 128 B (32 elements along M / N), 32 B chunk j of K row r stored at
// j ^ (r & 3).  LBO = stride between 128 B blocks along M / N, SBO = stride between groups of 4 K rows
// (cute::UMMA::make_umma_desc<Major::MN>).  With plain SWIZZLE_128B the tensor core returned zeros (measured).
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo >> 4) << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;
    return d;
}
// byte offset inside a 128 B K row of 16 B chunk c (0..7) of row r: its 32 B chunk c >> 1 moves to (c >> 1) ^ (r & 3)
__device__ __forceinline__ uint32_t swz_chunk(uint32_t c, uint32_t r) { return ((((c >> 1) ^ (r & 3u)) << 1) | (c & 1u)) << 4; }
// c_format F32 @4, a / b format TF32 (2) @7 / @10, a_major = b_major = MN (1) @15 / @16, N >> 3 @17, M >> 4 @24
__device__ __forceinline__ uint32_t make_idesc_mn(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// groups of this pass that have at least one offset present in the tile
__device__ __forceinline__ uint32_t live_groups(const Params& P, int tile, int g0, int gcnt, int kb_per_off) {
    if (!P.index || !P.tile_mask) return (1u << gcnt) - 1u;
    const uint32_t mask = __ldg(P.tile_mask + tile);
    uint32_t live = 0;
    for (int gi = 0; gi < gcnt; ++gi) {
        const int u0 = 4 * (g0 + gi), u1 = min(u0 + 3, P.units - 1);
        const int k0 = u0 / kb_per_off, k1 = u1 / kb_per_off;
        const uint32_t span = ((2u << k1) - 1u) & ~((1u << k0) - 1u);
        if (mask & span) live |= 1u << gi;
    }
    return live;
}

__global__ void __launch_bounds__(THREADS, 1) k_wgrad_tc(const Params P) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int N = P.c_out;
    Layout L;
    L.a0 = base;
    L.b_bytes = (uint32_t)(N / 32) * UNIT_BYTES;
    L.b0 = L.a0 + NA * A_STAGE;
    L.idx0 = L.b0 + NBS * L.b_bytes;
    L.bars = L.idx0 + idx_bytes(P.gp);
    L.tmem_slot = L.bars + BAR_BYTES;
    L.flags = L.tmem_slot + 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (L.tmem_slot - smem_u32(smem_raw)));
    volatile uint32_t* flags_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (L.flags - smem_u32(smem_raw)));

    const int kb_per_off = P.c_in / 32;
    const int g0 = (int)blockIdx.y * P.gp;
    const int gcnt = min(P.gp, P.groups - g0);
    const int t0 = (int)blockIdx.x * P.tiles_per_split;
    const int t1 = min(t0 + P.tiles_per_split, P.tiles);

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < NA; ++s) mbar_init(L.a_full(s), 128), mbar_init(L.a_empty(s), 1);
        for (uint32_t s = 0; s < NBS; ++s) mbar_init(L.b_full(s), 128), mbar_init(L.b_empty(s), 1);
        mbar_init(L.acc_full(), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(L.tmem_slot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp < 4) {
        // ===================== gather: warp w stages unit w of every group, all warps the d_out rows =====================
        const uint32_t c = (uint32_t)lane & 7u, rsub = (uint32_t)lane >> 3;     // 16 B chunk c of rows rsub + 4 i
        uint32_t a_slot = 0, a_phase = 0, b_slot = 0, b_phase = 0;
        // The rulebook rows this warp needs (offset of unit 4 g + warp, every group of the pass) are staged in a warp-private
        // shared-memory buffer one tile ahead: read straight from global memory, each stage paid a DRAM / L2 latency before
        // its copies could be issued (measured: 1 140 cycles per 16 KB stage, 317 us per level-0 launch of the 2-tile batch)
        const uint32_t ibuf0 = L.idx0 + (uint32_t)warp * 2u * (uint32_t)P.gp * (BM * 4u);
        auto stage_idx = [&](int tile, uint32_t buf) {
            if (P.index) {
                for (int gi = 0; gi < gcnt; ++gi) {
                    const int u = 4 * (g0 + gi) + warp;
                    if (u >= P.units) break;
                    const int32_t* ip = P.index + (int64_t)(u / kb_per_off) * P.index_stride + (int64_t)tile * BM + 4 * lane;
                    cp_async16(ibuf0 + (buf * (uint32_t)P.gp + (uint32_t)gi) * (BM * 4u) + 16u * (uint32_t)lane, ip, 16u);
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        uint32_t ibuf_sel = 0;
        if (t0 < t1) stage_idx(t0, 0u);
        for (int tile = t0; tile < t1; ++tile) {
            if (tile + 1 < t1) stage_idx(tile + 1, ibuf_sel ^ 1u);
            else asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 1;" ::: "memory");      // this tile's rows have landed (the next tile's may be in flight)
            __syncwarp();
            const uint32_t ibuf = ibuf0 + ibuf_sel * (uint32_t)P.gp * (BM * 4u);
            ibuf_sel ^= 1u;
            const uint32_t live = live_groups(P, tile, g0, gcnt, kb_per_off);
            if (!live) continue;
            for (int kbk = 0; kbk < BM / KR; ++kbk) {
                const int64_t rb = (int64_t)tile * BM + kbk * KR;
                if (rb >= P.n_out) break;
                // ---- d_out rows rb .. rb + 31: thread (warp, lane) copies chunk c of rows rsub + 4 i of blocks warp, warp + 4, ...
                mbar_wait(L.b_empty(b_slot), b_phase ^ 1u);
                for (int blk = warp; blk < N / 32; blk += 4) {
#pragma unroll
                    for (uint32_t i = 0; i < 8; ++i) {
                        const uint32_t r = rsub + 4u * i;
                        const int64_t row = rb + r;
                        const uint32_t dst = L.b(b_slot) + (uint32_t)blk * UNIT_BYTES + r * 128u + swz_chunk(c, r);
                        const float* sp = P.d_out + (row < P.n_out ? row : 0) * (int64_t)N + blk * 32 + c * 4;
                        cp_async16_cg(dst, sp, row < P.n_out ? 16u : 0u);
                    }
                }
                cp_async_mbar_arrive_noinc(L.b_full(b_slot));
                if (++b_slot == NBS) b_slot = 0, b_phase ^= 1u;
                // ---- one A stage per live group
                for (int gi = 0; gi < gcnt; ++gi) {
                    if (!((live >> gi) & 1u)) continue;
                    const int u = 4 * (g0 + gi) + warp;
                    mbar_wait(L.a_empty(a_slot), a_phase ^ 1u);
                    if (u < P.units) {
                        const int k = u / kb_per_off, kb = u - k * kb_per_off;
                        const float* sbase = P.src + kb * 32 + c * 4;
                        int64_t srow[8];
#pragma unroll
                        for (uint32_t i = 0; i < 8; ++i) {
                            const int64_t row = rb + rsub + 4u * i;
                            srow[i] = P.index ? (int64_t)ld_shared_i32(ibuf + (uint32_t)gi * (BM * 4u) + 4u * (uint32_t)(kbk * KR + (int)rsub + 4 * (int)i))
                                              : (row < P.n_out ? row : -1);
                        }
#pragma unroll
                        for (uint32_t i = 0; i < 8; ++i) {
                            const uint32_t r = rsub + 4u * i;
                            const uint32_t dst = L.a(a_slot) + (uint32_t)warp * UNIT_BYTES + r * 128u + swz_chunk(c, r);
                            cp_async16_cg(dst, sbase + (srow[i] >= 0 ? srow[i] : 0) * P.src_stride, srow[i] >= 0 ? 16u : 0u);
                        }
                    }
                    cp_async_mbar_arrive_noinc(L.a_full(a_slot));
                    if (++a_slot == NA) a_slot = 0, a_phase ^= 1u;
                }
            }
        }
    } else {
        // ===================== MMA issue =====================================================================================
        const uint32_t idesc = make_idesc_mn(N) ^ P.idesc_xor;
        uint32_t a_slot = 0, a_phase = 0, b_slot = 0, b_phase = 0, started = 0;
        for (int tile = t0; tile < t1; ++tile) {
            const uint32_t live = live_groups(P, tile, g0, gcnt, kb_per_off);
            if (!live) continue;
            for (int kbk = 0; kbk < BM / KR; ++kbk) {
                const int64_t rb = (int64_t)tile * BM + kbk * KR;
                if (rb >= P.n_out) break;
                mbar_wait(L.b_full(b_slot), b_phase);
                const uint64_t bd = make_desc_mn(L.b(b_slot), P.lbo_b, P.sbo);
                for (int gi = 0; gi < gcnt; ++gi) {
                    if (!((live >> gi) & 1u)) continue;
                    mbar_wait(L.a_full(a_slot), a_phase);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // cp.async (generic proxy) writes -> tensor core reads
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t ad = make_desc_mn(L.a(a_slot), P.lbo_a, P.sbo);
                        const uint32_t acc = tmem_base + (uint32_t)(gi * N);
#pragma unroll
                        for (uint32_t ks = 0; ks < KR / 8; ++ks)
                            umma_tf32(acc, ad + (uint64_t)(ks * 64u), bd + (uint64_t)(ks * 64u), idesc, (((started >> gi) & 1u) | ks) ? 1u : 0u);
                        umma_commit(L.a_empty(a_slot));
                    }
                    __syncwarp();
                    started |= 1u << gi;
                    if (++a_slot == NA) a_slot = 0, a_phase ^= 1u;
                }
                if (elect_one()) umma_commit(L.b_empty(b_slot));
                __syncwarp();
                if (++b_slot == NBS) b_slot = 0, b_phase ^= 1u;
            }
        }
        if (lane == 0) *flags_ptr = started;
        __syncwarp();
        if (elect_one()) umma_commit(L.acc_full());
        __syncwarp();
    }
    __syncthreads();     // `started` is visible; every gather has been issued

    if (warp < 4) {
        // ===================== write-out: warp w = unit w of every group, lane = input channel, 32 columns at a time ==========
        const uint32_t started = *flags_ptr;
        mbar_wait(L.acc_full(), 0u);
        tc_fence_after();
        const size_t dw_elems = (size_t)P.n_off * P.c_in * P.c_out;
        float* out = P.out + (size_t)blockIdx.x * dw_elems;
        for (int gi = 0; gi < gcnt; ++gi) {
            const int u = 4 * (g0 + gi) + warp;
            if (u >= P.units) continue;
            const int k = u / kb_per_off, kb = u - k * kb_per_off;
            float* dst = out + ((size_t)k * P.c_in + kb * 32 + lane) * N;
            for (int c0 = 0; c0 < N; c0 += 32) {
                uint32_t acc[32];
                if ((started >> gi) & 1u) {
                    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(gi * N + c0), acc);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[j] = 0u;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<uint4*>(dst + c0 + 4 * j) = make_uint4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// dW[e] = sum over the row splits, in split order (deterministic)
__global__ void __launch_bounds__(256) k_wgrad_reduce(const float* __restrict__ part, int splits, int64_t elems4, float* __restrict__ dw) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= elems4) return;
    float4 s = __ldg(reinterpret_cast<const float4*>(part) + i);
    for (int p = 1; p < splits; ++p) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(part) + (int64_t)p * elems4 + i);
        s.x += v.x, s.y += v.y, s.z += v.z, s.w += v.w;
    }
    reinterpret_cast<float4*>(dw)[i] = s;
}

struct Plan {
    int units, groups, gp, passes, tiles, splits, tiles_per_split;
};
static Plan make_plan(int64_t n_out, int c_in, int n_off, int c_out) {
    Plan p;
    p.units = n_off * (c_in / 32);
    p.groups = (p.units + 3) / 4;
    p.gp = TMEM_COLS / c_out;
    if (p.gp > p.groups) p.gp = p.groups;
    if (p.gp > 16) p.gp = 16;
    p.passes = (p.groups + p.gp - 1) / p.gp;
    p.tiles = (int)((n_out + BM - 1) / BM);
    const size_t dw_bytes = (size_t)n_off * c_in * c_out * sizeof(float);
    int want = (148 + p.passes - 1) / p.passes;                  // about one wave of CTAs
    const size_t cap = (size_t)96 << 20;                          // partial-sum scratch
    if ((size_t)want * dw_bytes > cap) want = (int)(cap / dw_bytes);
    if (want < 1) want = 1;
    if (want > p.tiles) want = p.tiles;
    p.tiles_per_split = (p.tiles + want - 1) / want;
    p.splits = (p.tiles + p.tiles_per_split - 1) / p.tiles_per_split;
    return p;
}

}  // namespace wg
}  // namespace tl

using namespace tl;

extern "C" {

int tl_conv_wgrad_tc_eligible(int32_t c_in, int32_t c_out) { return c_in % 32 == 0 && c_in >= 32 && c_out % 32 == 0 && c_out >= 32 && c_out <= 256; }

size_t tl_conv_wgrad_tc_workspace_bytes(int64_t n_out, int32_t c_in, int32_t n_off, int32_t c_out) {
    if (n_out <= 0 || !tl_conv_wgrad_tc_eligible(c_in, c_out)) return 256;
    const wg::Plan p = wg::make_plan(n_out, c_in, n_off, c_out);
    return p.splits > 1 ? (size_t)p.splits * n_off * c_in * c_out * sizeof(float) + 256 : 256;
}

// Same contract as tl_conv_wgrad (dw [n_off][c_in][c_out] fp32, overwritten) for tcgen05-eligible widths; TF32 operands.
int tl_conv_wgrad_tc(const float* src, int64_t src_stride, int32_t c_in, int32_t n_off, const int32_t* index, int64_t index_stride,
                     const uint32_t* tile_mask, const float* d_out, int64_t n_out, int32_t c_out, float* dw, void* workspace,
                     size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    TL_REQUIRE(src && d_out && dw && n_off >= 1 && n_off <= 27, "tl_conv_wgrad_tc: bad arguments");
    TL_REQUIRE(tl_conv_wgrad_tc_eligible(c_in, c_out), "tl_conv_wgrad_tc: c_in=%d c_out=%d not eligible (multiples of 32, c_out <= 256)", c_in, c_out);
    TL_REQUIRE(index || n_off == 1, "tl_conv_wgrad_tc: identity map needs n_off == 1");
    TL_REQUIRE(src_stride % 4 == 0, "tl_conv_wgrad_tc: src_stride must be a multiple of 4");
    const size_t dw_elems = (size_t)n_off * c_in * c_out;
    if (n_out <= 0) {
        TL_CUDA_CHECK(cudaMemsetAsync(dw, 0, sizeof(float) * dw_elems, stream));
        return TL_OK;
    }
    const wg::Plan pl = wg::make_plan(n_out, c_in, n_off, c_out);
    TL_REQUIRE(pl.splits == 1 || (workspace && workspace_bytes >= (size_t)pl.splits * dw_elems * sizeof(float)),
               "tl_conv_wgrad_tc: workspace too small (%zu bytes)", workspace_bytes);
    wg::Params P;
    P.src = src, P.src_stride = src_stride, P.c_in = c_in, P.n_off = n_off, P.index = index, P.index_stride = index_stride;
    P.tile_mask = tile_mask, P.d_out = d_out, P.n_out = n_out, P.c_out = c_out;
    P.out = pl.splits == 1 ? dw : reinterpret_cast<float*>(workspace);
    P.units = pl.units, P.groups = pl.groups, P.gp = pl.gp, P.tiles = pl.tiles, P.tiles_per_split = pl.tiles_per_split;
    auto env_u32 = [](const char* name, uint32_t dflt) { const char* v = getenv(name); return v ? (uint32_t)strtoul(v, nullptr, 0) : dflt; };
    P.lbo_a = env_u32("TL_WG_LBO_A", wg::UNIT_BYTES), P.lbo_b = env_u32("TL_WG_LBO_B", wg::UNIT_BYTES), P.sbo = env_u32("TL_WG_SBO", 512u);
    P.idesc_xor = env_u32("TL_WG_IDESC_XOR", 0u);
    static bool configured[16] = {false};
    int dev = 0;
    TL_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 16 && !configured[dev]) {      // function attributes are per device
        TL_CUDA_CHECK(cudaFuncSetAttribute(wg::k_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured[dev] = true;
    }
    size_t smem = wg::smem_bytes(c_out, pl.gp);
    if (smem < 120 * 1024) smem = 120 * 1024;      // one CTA per SM: each allocates all 512 tensor-memory columns
    TL_REQUIRE(smem <= 227 * 1024, "tl_conv_wgrad_tc: shared memory %zu", smem);
    wg::k_wgrad_tc<<<dim3((unsigned)pl.splits, (unsigned)pl.passes), wg::THREADS, smem, stream>>>(P);
    TL_LAUNCH_CHECK();
    if (pl.splits > 1) {
        const int64_t e4 = (int64_t)(dw_elems / 4);
        wg::k_wgrad_reduce<<<(unsigned)((e4 + 255) / 256), 256, 0, stream>>>(P.out, pl.splits, e4, dw);
        TL_LAUNCH_CHECK();
    }
    return TL_OK;
}

}  // extern "C"
