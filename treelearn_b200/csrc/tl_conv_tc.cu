// tcgen05 / TMEM segmented gather-GEMM sparse convolution (mode TL_MODE_TF32), sm_100a only.
//
// One CTA owns a tile of 128 output voxels (UMMA M=128, cta_group::1) and all C_out columns
// (UMMA N = C_out, fp32 accumulators in TMEM, double-buffered so the epilogue of tile i overlaps the
// main loop of tile i+1).  The K loop runs over (segment, kernel offset, 32-channel block): per step
//   * 4 producer warps gather the 128 neighbour rows (128 B each, cp.async 16 B, zero-fill for absent
//     neighbours) into a 128B-swizzled K-major A stage and copy the offset's [C_out x 32] weight slab into
//     a B stage -- exactly the canonical SWIZZLE_128B layouts the UMMA shared-memory descriptors name;
//   * 1 MMA thread issues 4 x tcgen05.mma.kind::tf32 (K=8 each) and tcgen05.commit's the stage back;
//   * 4 epilogue warps tcgen05.ld the accumulator rows, add the residual, and write up to three outputs
//     (raw, and relu(scale*v+shift) for the next layers' BatchNorm+ReLU, rounded to TF32 so the tensor
//     core's operand truncation of those tensors is exact).
// Offsets that no voxel of the tile uses are skipped through the rulebook's per-tile bitmask.
// Weights arrive pre-rounded (RN) to TF32, layout [n_off][C_out][C_in] (K-major B operand).
#include "tl_common.cuh"

namespace tl {
namespace tc {

constexpr int BM = 128;          // rows per tile == TMEM lanes
constexpr int BK = 32;           // fp32 elements per K block == one 128 B swizzle row
constexpr int STAGES = 4;
constexpr int LAG = 2;           // cp.async groups in flight per producer thread before the oldest is published
constexpr int A_STAGE_BYTES = BM * 128;
constexpr int kProducerThreads = 128, kEpilogueThreads = 128;
constexpr int kThreads = kProducerThreads + kEpilogueThreads + 32;

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
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
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

struct Layout {  // dynamic shared memory carve-up (1024 B aligned base)
    uint32_t a[STAGES], b[STAGES];
    uint32_t full[STAGES], empty[STAGES], tfull[2], tempty[2];
    uint32_t tmem_slot;
};

__device__ __forceinline__ Layout carve(uint32_t base, int n) {
    Layout L;
    uint32_t off = base;
#pragma unroll
    for (int s = 0; s < STAGES; ++s) L.a[s] = off + s * A_STAGE_BYTES;
    off += STAGES * A_STAGE_BYTES;
#pragma unroll
    for (int s = 0; s < STAGES; ++s) L.b[s] = off + s * (n * 128);
    off += STAGES * (n * 128);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) L.full[s] = off + 8 * s, L.empty[s] = off + 8 * (STAGES + s);
    off += 16 * STAGES;
    L.tfull[0] = off, L.tfull[1] = off + 8, L.tempty[0] = off + 16, L.tempty[1] = off + 24;
    L.tmem_slot = off + 32;
    return L;
}
static inline size_t smem_bytes(int n) { return 1024 + STAGES * (A_STAGE_BYTES + (size_t)n * 128) + 16 * STAGES + 64; }

__device__ __forceinline__ uint32_t seg_mask(const tl_conv_seg& sg, int64_t tile) {
    if (!sg.index) return 1u;
    const uint32_t all = sg.n_off >= 32 ? 0xffffffffu : ((1u << sg.n_off) - 1u);
    return (sg.tile_mask ? sg.tile_mask[tile] : 0xffffffffu) & all;
}

__global__ void __launch_bounds__(kThreads, 1) k_conv_tc(const tl_conv_desc d, int num_tiles, int tmem_cols, int buf_cols) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int N = d.c_out;
    const Layout L = carve(base, N);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (L.tmem_slot - smem_u32(smem_raw)));

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(L.full[s], kProducerThreads);
            mbar_init(L.empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(L.tfull[b], 1);
            mbar_init(L.tempty[b], kEpilogueThreads);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {  // TMEM allocation is owned by the MMA warp
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(L.tmem_slot), "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp < 4) {
        // ===================== producers: gather A rows + copy B slab ==============================
        const int tid = threadIdx.x;
        uint32_t it = 0;      // chunks issued
        uint32_t pub = 0;     // chunks published (arrived on full[])
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int64_t row0 = (int64_t)tile * BM;
            for (int s = 0; s < d.n_seg; ++s) {
                const tl_conv_seg sg = d.seg[s];
                const uint32_t mask = seg_mask(sg, tile);
                const int kblocks = sg.c_in / BK;
                for (int k = 0; k < sg.n_off; ++k) {
                    if (!((mask >> k) & 1u)) continue;
                    int my_row;
                    {
                        const int64_t r = row0 + warp * 32 + lane;
                        if (sg.index) my_row = sg.index[(int64_t)k * sg.index_stride + r];
                        else my_row = r < d.n_out ? (int)r : -1;
                    }
                    const float* wk = sg.weight + (int64_t)k * N * sg.c_in;
                    for (int kb = 0; kb < kblocks; ++kb) {
                        const uint32_t slot = it % STAGES;
                        mbar_wait(L.empty[slot], ((it / STAGES) & 1u) ^ 1u);
                        const uint32_t a_st = L.a[slot], b_st = L.b[slot];
                        const int chunk = lane & 7;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int rl = i * 4 + (lane >> 3);
                            const int r = __shfl_sync(0xffffffffu, my_row, rl);
                            const float* src = sg.src + (r >= 0 ? (int64_t)r * sg.src_stride + kb * BK + chunk * 4 : 0);
                            cp_async16(a_st + (warp * 32 + rl) * 128 + ((chunk ^ (rl & 7)) << 4), src, r >= 0 ? 16u : 0u);
                        }
                        for (int e = tid; e < N * 8; e += kProducerThreads) {
                            const int n = e >> 3, c = e & 7;
                            cp_async16(b_st + n * 128 + ((c ^ (n & 7)) << 4), wk + (int64_t)n * sg.c_in + kb * BK + c * 4, 16u);
                        }
                        cp_async_commit();
                        ++it;
                        if (it - pub > LAG) {
                            cp_async_wait<LAG>();
                            fence_proxy_async();
                            mbar_arrive(L.full[pub % STAGES]);
                            ++pub;
                        }
                    }
                }
            }
        }
        cp_async_wait<0>();
        fence_proxy_async();
        for (; pub < it; ++pub) mbar_arrive(L.full[pub % STAGES]);
    } else if (warp < 8) {
        // ===================== epilogue: TMEM -> registers -> global ================================
        const int ew = warp - 4;  // == warp % 4: the TMEM lane quarter this warp may touch
        uint32_t titer = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++titer) {
            const uint32_t buf = titer & 1u;
            bool any = false;
            for (int s = 0; s < d.n_seg; ++s) any = any || (seg_mask(d.seg[s], tile) != 0u);
            mbar_wait(L.tfull[buf], (titer >> 1) & 1u);
            tc_fence_after();
            const int64_t row = (int64_t)tile * BM + ew * 32 + lane;
            const bool live = row < d.n_out;
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + buf * buf_cols;
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
                        for (int j = 0; j < 8; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    }
                    if (d.out_act1) {
                        float4* op = reinterpret_cast<float4*>(d.out_act1 + o);
                        const float4* sp = reinterpret_cast<const float4*>(d.scale1 + c0);
                        const float4* tp = reinterpret_cast<const float4*>(d.shift1 + c0);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 s4 = __ldg(sp + j), t4 = __ldg(tp + j);
                            op[j] = make_float4(round_tf32(fmaxf(fmaf(v[4 * j], s4.x, t4.x), 0.f)),
                                                round_tf32(fmaxf(fmaf(v[4 * j + 1], s4.y, t4.y), 0.f)),
                                                round_tf32(fmaxf(fmaf(v[4 * j + 2], s4.z, t4.z), 0.f)),
                                                round_tf32(fmaxf(fmaf(v[4 * j + 3], s4.w, t4.w), 0.f)));
                        }
                    }
                    if (d.out_act2) {
                        float4* op = reinterpret_cast<float4*>(d.out_act2 + o);
                        const float4* sp = reinterpret_cast<const float4*>(d.scale2 + c0);
                        const float4* tp = reinterpret_cast<const float4*>(d.shift2 + c0);
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
            tc_fence_before();
            mbar_arrive(L.tempty[buf]);
        }
    } else {
        // ===================== MMA issuer (one elected thread) ======================================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(N);
            uint32_t it = 0, titer = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++titer) {
                const uint32_t buf = titer & 1u;
                mbar_wait(L.tempty[buf], ((titer >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + buf * buf_cols;
                uint32_t first = 1;
                for (int s = 0; s < d.n_seg; ++s) {
                    const tl_conv_seg& sg = d.seg[s];
                    const uint32_t mask = seg_mask(sg, tile);
                    const int kblocks = sg.c_in / BK;
                    for (int k = 0; k < sg.n_off; ++k) {
                        if (!((mask >> k) & 1u)) continue;
                        for (int kb = 0; kb < kblocks; ++kb) {
                            const uint32_t slot = it % STAGES;
                            mbar_wait(L.full[slot], (it / STAGES) & 1u);
                            tc_fence_after();
                            const uint64_t adesc = make_smem_desc(L.a[slot]);
                            const uint64_t bdesc = make_smem_desc(L.b[slot]);
#pragma unroll
                            for (int kk = 0; kk < BK / 8; ++kk) {  // UMMA_K = 8 for tf32: advance 32 B inside the atom
                                umma_tf32(tmem_d, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc,
                                          (first && kk == 0) ? 0u : 1u);
                            }
                            first = 0;
                            umma_commit(L.empty[slot]);
                            ++it;
                        }
                    }
                }
                if (first) mbar_arrive(L.tfull[buf]);  // no pair in this tile: epilogue writes zeros
                else umma_commit(L.tfull[buf]);
            }
        }
        __syncwarp();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
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

int conv_fwd_tc(const tl_conv_desc& d, cudaStream_t stream) {
    if (!tc_eligible(d)) return conv_fwd_simt(d, stream);  // e.g. the 4-channel input conv: K < one UMMA K block
    static int num_sms = 0;
    if (!num_sms) {
        int dev = 0;
        TL_CUDA_CHECK(cudaGetDevice(&dev));
        TL_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        TL_CUDA_CHECK(cudaFuncSetAttribute(tc::k_conv_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    const int n = d.c_out;
    int buf_cols = 32;
    while (buf_cols < n) buf_cols <<= 1;
    const int tmem_cols = 2 * buf_cols;  // <= 512
    const size_t smem = tc::smem_bytes(n);
    TL_REQUIRE(smem <= 227 * 1024, "tl_conv_fwd(tf32): c_out=%d needs %zu B shared memory", n, smem);
    const int num_tiles = (d.n_out + tc::BM - 1) / tc::BM;
    // CTAs per SM bounded by shared memory and by TMEM columns (512 per SM)
    int per_sm = (int)((227 * 1024) / smem);
    if (per_sm > 512 / tmem_cols) per_sm = 512 / tmem_cols;
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    int grid = num_sms * per_sm;
    if (grid > num_tiles) grid = num_tiles;
    tc::k_conv_tc<<<grid, tc::kThreads, smem, stream>>>(d, num_tiles, tmem_cols, buf_cols);
    TL_LAUNCH_CHECK();
    return TL_OK;
}

}  // namespace tl
