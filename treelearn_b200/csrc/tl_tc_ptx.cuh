// PTX wrappers shared by the tcgen05 convolution kernels (tl_conv_tc.cu: A operand in shared memory; tl_conv_ts.cu: A in TMEM).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace tl {
namespace tc {

constexpr int BM = 128;          // rows per tile == TMEM lanes

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
// per-tile waits (descriptor / accumulator hand-offs) back off between polls so that they do not steal issue slots
// from the gather warps (measured: 20 % of all issued instructions were try_wait spins)
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity, uint32_t ns) {
    uint32_t done = 0;
    while (true) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) break;
        if (ns) __nanosleep(ns);
    }
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async16_cg(uint32_t dst, const void* src, uint32_t src_bytes) {   // L2 only
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ int ld_shared_i32(uint32_t addr) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
// the mbarrier is arrived on (without bumping its pending count) once all prior cp.async of this thread completed
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 (ignored for swizzled K-major, set 1) | [32,46) SBO >> 4 (8 rows
//   x 128 B = 1024) | [46,48) version = 1 (sm100) | [49,52) base offset = 0 (1024 B aligned stages) |
//   [61,64) layout = 2 (SWIZZLE_128B)
// row_bytes 128 -> SWIZZLE_128B (layout 2, SBO 1024); row_bytes 64 -> SWIZZLE_64B (layout 4, SBO 512)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, int row_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((8 * row_bytes) >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(row_bytes == 128 ? 2 : 4) << 61;
    return d;
}
// cute::UMMA::InstrDescriptor: c_format F32 (1) @4, a/b format TF32 (2) @7/@10, K-major both, N>>3 @17, M>>4 @24
__device__ __forceinline__ uint32_t make_idesc(int n, bool half) {   // a/b format: 2 = TF32, 0 = F16
    const uint32_t fmt = half ? 0u : 2u;
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
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
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// accumulate-always forms (enable_input_d = true folds to the constant predicate: no setp / predicate moves per MMA)
__device__ __forceinline__ void umma_tf32_acc(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.eq.u32 p, 1, 1;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc)
        : "memory");
}
__device__ __forceinline__ void umma_f16_acc(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.eq.u32 p, 1, 1;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc)
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

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}
// fp32 pair -> (hi, lo) fp16 pairs: hi = rn(v), lo = rn(v - hi)  (two-term split, |v - hi - lo| <= 2^-22 |v|)
__device__ __forceinline__ void split_half2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half ha = __float2half_rn(a), hb = __float2half_rn(b);
    const __half2 h = __halves2half2(ha, hb);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = pack_half2(a - __half2float(ha), b - __half2float(hb));
}

// Operand formats of the activated tensors the next convolution gathers:
//   FMT_TF32  [row][C] fp32 rounded to TF32            (TL_MODE_TF32)
//   FMT_F16   [row][C] fp16                            (TL_MODE_F16)
//   FMT_F16X2 [row][C/32][2][32] fp16: per 32-channel block 64 B of hi halves, then 64 B of lo halves (TL_MODE_F16X2);
//             a gathered 32-channel row piece stays one contiguous 128 B line
constexpr int FMT_TF32 = 4, FMT_F16 = 2, FMT_F16X2 = 22;
// P-layout of the tensor-memory-A kernel (tl_conv_ts.cu): logical channel stored at position m of a 32-channel block
__host__ __device__ constexpr int p_chan(int m) { return 8 * ((m % 8) / 2) + 2 * (m / 8) + (m % 2); }

// 4 consecutive activated columns (col % 4 == 0) of one row -> the consumer's operand format
template <int FMT>
__device__ __forceinline__ void store_act4(void* out, int64_t row, int n, int col, float a0, float a1, float a2, float a3) {
    if (FMT == FMT_TF32) {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + row * n + col) =
            make_float4(round_tf32(a0), round_tf32(a1), round_tf32(a2), round_tf32(a3));
    } else if (FMT == FMT_F16) {
        *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(out) + row * n + col) = make_uint2(pack_half2(a0, a1), pack_half2(a2, a3));
    } else {
        uint32_t h0, l0, h1, l1;
        split_half2(a0, a1, h0, l0);
        split_half2(a2, a3, h1, l1);
        char* p = reinterpret_cast<char*>(out) + (row * n + (col & ~31)) * 4 + (col & 31) * 2;
        *reinterpret_cast<uint2*>(p) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(p + 64) = make_uint2(l0, l1);
    }
}
template <int FMT>
__device__ __forceinline__ void store_act1(void* out, int64_t e, int n, float a) {
    if (FMT == FMT_TF32) {
        reinterpret_cast<float*>(out)[e] = round_tf32(a);
    } else if (FMT == FMT_F16) {
        reinterpret_cast<__half*>(out)[e] = __float2half_rn(a);
    } else {
        const int64_t row = e / n;
        const int col = (int)(e - row * n);
        const __half h = __float2half_rn(a);
        __half* p = reinterpret_cast<__half*>(reinterpret_cast<char*>(out) + (row * n + (col & ~31)) * 4) + (col & 31);
        p[0] = h;
        p[32] = __float2half_rn(a - __half2float(h));
    }
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// TMA bulk copy global -> shared, completion (bytes) signalled on the mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 ld_shared_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint2 ld_shared_u2(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}

// one lane of a converged warp (cute::elect_one_sync): keeps the surrounding loop warp-uniform, so the compiler holds
// descriptors / barrier addresses in uniform registers instead of wrapping every tcgen05 / bulk-copy instruction in a
// per-lane "waterfall" loop (measured: ~375 issue cycles per chunk in the old `if (lane == 0)` MMA loop)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n"
        ".reg .b32 rx;\n"
        ".reg .pred px;\n"
        "elect.sync rx|px, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, px;\n"
        "}\n"
        : "=r"(pred)::"memory");
    return pred != 0;
}

}  // namespace tc
}  // namespace tl
