// fp32 SIMT segmented gather-GEMM sparse convolution (mode TL_MODE_FP32): the exact-arithmetic
// ("parity") path and the fallback for channel widths the tcgen05 path does not take.
// One CTA = TM output rows x all C_out columns; per kernel offset the gathered input rows and
// the offset's weight slab are staged in shared memory in 32-channel chunks; each thread keeps an
// RPT x NCOL register tile (rows warp-striped, columns lane-striped => coalesced epilogue).
#include <cuda_fp16.h>
#include <stdlib.h>

#include "tl_common.cuh"
#include "tl_tc_ptx.cuh"

namespace tl {

constexpr int kKC = 32;        // channels per staged chunk
constexpr int kAStride = 36;   // floats; 144 B rows keep float4 alignment, broadcast reads => no conflicts

template <int NCOL, int RPT>
__global__ void __launch_bounds__(256) k_conv_simt(const tl_conv_desc d) {
    constexpr int TM = 8 * RPT;
    __shared__ __align__(16) float As[TM * kAStride];
    __shared__ __align__(16) float Ws[kKC * NCOL * 32];
    __shared__ int s_rows[TM];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t row0 = (int64_t)blockIdx.x * TM;
    const int c_out = d.c_out;

    float acc[RPT][NCOL];
#pragma unroll
    for (int i = 0; i < RPT; ++i)
#pragma unroll
        for (int j = 0; j < NCOL; ++j) acc[i][j] = 0.f;

    for (int s = 0; s < d.n_seg; ++s) {
        const tl_conv_seg sg = d.seg[s];
        const uint32_t tmask = sg.tile_mask ? sg.tile_mask[row0 / TL_TILE_ROWS] : 0xffffffffu;
        for (int k = 0; k < sg.n_off; ++k) {
            if (sg.index && !((tmask >> k) & 1u)) continue;  // block-uniform
            __syncthreads();
            if (tid < TM) {
                const int64_t r = row0 + tid;
                s_rows[tid] = (r < d.n_out) ? (sg.index ? sg.index[(int64_t)k * sg.index_stride + r] : (int)r) : -1;
            }
            __syncthreads();
            for (int c0 = 0; c0 < sg.c_in; c0 += kKC) {
#pragma unroll
                for (int it = 0; it < TM * 8 / 256; ++it) {
                    const int slot = tid + it * 256;
                    const int rr = slot >> 3, cc = (slot & 7) * 4;
                    const int srow = s_rows[rr];
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (srow >= 0 && c0 + cc < sg.c_in)
                        v = __ldg(reinterpret_cast<const float4*>(sg.src + (int64_t)srow * sg.src_stride + c0 + cc));
                    *reinterpret_cast<float4*>(&As[rr * kAStride + cc]) = v;
                }
                const float* wk = sg.weight + ((int64_t)k * sg.c_in + c0) * c_out;
                for (int e = tid; e < kKC * NCOL * 32; e += 256) {
                    const int kk = e / (NCOL * 32), col = e % (NCOL * 32);
                    Ws[e] = (c0 + kk < sg.c_in && col < c_out) ? __ldg(wk + (int64_t)kk * c_out + col) : 0.f;
                }
                __syncthreads();
#pragma unroll 2
                for (int kk4 = 0; kk4 < kKC; kk4 += 4) {
                    float b[4][NCOL];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int j = 0; j < NCOL; ++j) b[u][j] = Ws[(kk4 + u) * (NCOL * 32) + lane + 32 * j];
#pragma unroll
                    for (int i = 0; i < RPT; ++i) {
                        const float4 a = *reinterpret_cast<const float4*>(&As[(warp * RPT + i) * kAStride + kk4]);
#pragma unroll
                        for (int j = 0; j < NCOL; ++j) {
                            acc[i][j] = fmaf(a.x, b[0][j], acc[i][j]);
                            acc[i][j] = fmaf(a.y, b[1][j], acc[i][j]);
                            acc[i][j] = fmaf(a.z, b[2][j], acc[i][j]);
                            acc[i][j] = fmaf(a.w, b[3][j], acc[i][j]);
                        }
                    }
                }
                __syncthreads();
            }
        }
    }

#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        const int64_t r = row0 + warp * RPT + i;
        if (r >= d.n_out) continue;
#pragma unroll
        for (int j = 0; j < NCOL; ++j) {
            const int col = lane + 32 * j;
            if (col >= c_out) continue;
            const int64_t o = r * c_out + col;
            float v = acc[i][j];
            if (d.residual) v += __ldg(d.residual + o);
            if (d.out_raw) d.out_raw[o] = v;
            if (d.out_act1) d.out_act1[o] = fmaxf(fmaf(v, __ldg(d.scale1 + col), __ldg(d.shift1 + col)), 0.f);
            if (d.out_act2) d.out_act2[o] = fmaxf(fmaf(v, __ldg(d.scale2 + col), __ldg(d.shift2 + col)), 0.f);
        }
    }
}

template <int NCOL, int RPT>
static int launch_simt(const tl_conv_desc& d, cudaStream_t stream) {
    constexpr int TM = 8 * RPT;
    const unsigned grid = (unsigned)((d.n_out + TM - 1) / TM);
    k_conv_simt<NCOL, RPT><<<grid, 256, 0, stream>>>(d);
    TL_LAUNCH_CHECK();
    return TL_OK;
}

int conv_fwd_simt(const tl_conv_desc& d, cudaStream_t stream) {
    switch ((d.c_out + 31) / 32) {
        case 1: return launch_simt<1, 16>(d, stream);
        case 2: return launch_simt<2, 16>(d, stream);
        case 3: return launch_simt<3, 16>(d, stream);
        case 4: return launch_simt<4, 16>(d, stream);
        case 5: return launch_simt<5, 8>(d, stream);
        case 6: return launch_simt<6, 8>(d, stream);
        case 7: return launch_simt<7, 8>(d, stream);
        case 8: return launch_simt<8, 8>(d, stream);
    }
    set_error("tl_conv_fwd(fp32): c_out=%d > 256 unsupported", d.c_out);
    return TL_ERR_UNSUPPORTED;
}

int conv_fwd_tc(const tl_conv_desc& d, cudaStream_t stream, bool half);  // tl_conv_tc.cu
int conv_fwd_in4(const tl_conv_desc& d, cudaStream_t stream, int fmt, bool perm);   // tl_conv_tc.cu
bool conv_ts_eligible(const tl_conv_desc& d);                            // tl_conv_ts.cu
int conv_fwd_ts(const tl_conv_desc& d, cudaStream_t stream, int nsplit, uint32_t src_fp32_mask);
bool conv_grp_eligible(const tl_conv_desc& d, uint32_t src_fp32_mask);   // tl_conv_grp.cu
int conv_fwd_grp(const tl_conv_desc& d, cudaStream_t stream, int nsplit);
bool conv_halo_eligible(const tl_conv_desc& d, uint32_t src_fp32_mask);  // tl_conv_halo.cu
int conv_fwd_halo(const tl_conv_desc& d, cudaStream_t stream, int nsplit);

// ------------------------------------------------------------------------------------------------
// voxel -> point gather + both MLP heads, one thread per point, weights broadcast from smem
// ------------------------------------------------------------------------------------------------
template <int C, int FMT>   // FMT: 0 fp32 rows, 1 fp16 rows, 2 f16x2 operand format ([C/32][2][32] fp16: hi + lo, P-layout), 3 fp16 rows in P-layout
__global__ void __launch_bounds__(128) k_heads(const void* __restrict__ vfeat_, const int64_t* __restrict__ v2p,
                                               int64_t n, const float* __restrict__ sw1, const float* __restrict__ sb1,
                                               const float* __restrict__ sw2, const float* __restrict__ sb2,
                                               const float* __restrict__ ow1, const float* __restrict__ ob1,
                                               const float* __restrict__ ow2, const float* __restrict__ ob2,
                                               float* __restrict__ feats, float* __restrict__ logits,
                                               float* __restrict__ offsets) {
    __shared__ __align__(16) float w1[2][C * C];
    __shared__ float b1[2][C];
    __shared__ float w2[5][C];
    __shared__ float b2[5];
    for (int e = threadIdx.x; e < C * C; e += blockDim.x) {
        w1[0][e] = sw1[e];
        w1[1][e] = ow1[e];
    }
    for (int e = threadIdx.x; e < C; e += blockDim.x) {
        b1[0][e] = sb1[e];
        b1[1][e] = ob1[e];
        w2[0][e] = sw2[e];
        w2[1][e] = sw2[C + e];
        w2[2][e] = ow2[e];
        w2[3][e] = ow2[C + e];
        w2[4][e] = ow2[2 * C + e];
    }
    if (threadIdx.x < 2) b2[threadIdx.x] = sb2[threadIdx.x];
    if (threadIdx.x >= 2 && threadIdx.x < 5) b2[threadIdx.x] = ob2[threadIdx.x - 2];
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4* dst = reinterpret_cast<float4*>(feats + i * C);
    float x[C];
    if (FMT == 2) {
        const uint2* src = reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(vfeat_) + v2p[i] * (2 * C));
#pragma unroll
        for (int c = 0; c < C / 4; ++c) {
            const int blk = (4 * c) / 32, j = (4 * c) % 32;      // 16 uint2 per 64 B plane, hi plane then lo plane per block
            const uint2 rh = __ldg(src + blk * 16 + j / 4), rl = __ldg(src + blk * 16 + 8 + j / 4);
            const float2 h0 = __half22float2(*reinterpret_cast<const __half2*>(&rh.x)), h1 = __half22float2(*reinterpret_cast<const __half2*>(&rh.y));
            const float2 l0 = __half22float2(*reinterpret_cast<const __half2*>(&rl.x)), l1 = __half22float2(*reinterpret_cast<const __half2*>(&rl.y));
            // positions 4c .. 4c+3 of the P-layout row -> logical channels
            x[blk * 32 + tc::p_chan(j)] = h0.x + l0.x, x[blk * 32 + tc::p_chan(j + 1)] = h0.y + l0.y;
            x[blk * 32 + tc::p_chan(j + 2)] = h1.x + l1.x, x[blk * 32 + tc::p_chan(j + 3)] = h1.y + l1.y;
        }
#pragma unroll
        for (int c = 0; c < C / 4; ++c) dst[c] = make_float4(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]);
    } else if (FMT == 3) {
        const uint2* src = reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(vfeat_) + v2p[i] * C);
#pragma unroll
        for (int c = 0; c < C / 4; ++c) {
            const int blk = (4 * c) / 32, j = (4 * c) % 32;
            const uint2 raw = __ldg(src + c);
            const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
            const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
            x[blk * 32 + tc::p_chan(j)] = lo.x, x[blk * 32 + tc::p_chan(j + 1)] = lo.y;
            x[blk * 32 + tc::p_chan(j + 2)] = hi.x, x[blk * 32 + tc::p_chan(j + 3)] = hi.y;
        }
#pragma unroll
        for (int c = 0; c < C / 4; ++c) dst[c] = make_float4(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]);
    } else if (FMT == 1) {
        const uint2* src = reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(vfeat_) + v2p[i] * C);
#pragma unroll
        for (int c = 0; c < C / 4; ++c) {
            const uint2 raw = __ldg(src + c);
            const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
            const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
            x[4 * c] = lo.x, x[4 * c + 1] = lo.y, x[4 * c + 2] = hi.x, x[4 * c + 3] = hi.y;
            dst[c] = make_float4(lo.x, lo.y, hi.x, hi.y);
        }
    } else {
        const float4* src = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(vfeat_) + v2p[i] * C);
#pragma unroll
        for (int c = 0; c < C / 4; ++c) {
            const float4 v = __ldg(src + c);
            dst[c] = v;
            x[4 * c] = v.x, x[4 * c + 1] = v.y, x[4 * c + 2] = v.z, x[4 * c + 3] = v.w;
        }
    }
    float y[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) y[q] = b2[q];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll 4
        for (int o = 0; o < C; ++o) {
            float a = b1[h][o];
            const float4* wr = reinterpret_cast<const float4*>(&w1[h][o * C]);
#pragma unroll
            for (int c = 0; c < C / 4; ++c) {
                const float4 w = wr[c];
                a = fmaf(w.x, x[4 * c], a);
                a = fmaf(w.y, x[4 * c + 1], a);
                a = fmaf(w.z, x[4 * c + 2], a);
                a = fmaf(w.w, x[4 * c + 3], a);
            }
            a = fmaxf(a, 0.f);
            if (h == 0) {
                y[0] = fmaf(w2[0][o], a, y[0]);
                y[1] = fmaf(w2[1][o], a, y[1]);
            } else {
                y[2] = fmaf(w2[2][o], a, y[2]);
                y[3] = fmaf(w2[3][o], a, y[3]);
                y[4] = fmaf(w2[4][o], a, y[4]);
            }
        }
    }
    logits[i * 2] = y[0];
    logits[i * 2 + 1] = y[1];
    offsets[i * 3] = y[2];
    offsets[i * 3 + 1] = y[3];
    offsets[i * 3 + 2] = y[4];
}

}  // namespace tl

using namespace tl;

extern "C" {

int tl_conv_fwd(const tl_conv_desc* desc, int32_t mode, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    TL_REQUIRE(desc != nullptr, "tl_conv_fwd: null descriptor");
    const tl_conv_desc& d = *desc;
    if (d.n_out == 0) return TL_OK;
    TL_REQUIRE(d.n_out > 0 && d.c_out > 0, "tl_conv_fwd: n_out=%d c_out=%d", d.n_out, d.c_out);
    TL_REQUIRE(d.n_seg >= 1 && d.n_seg <= TL_MAX_SEG, "tl_conv_fwd: n_seg=%d", d.n_seg);
    for (int s = 0; s < d.n_seg; ++s) {
        const tl_conv_seg& g = d.seg[s];
        TL_REQUIRE(g.src && g.weight && g.c_in > 0 && g.c_in % 4 == 0 && g.src_stride % 4 == 0,
                   "tl_conv_fwd: segment %d: c_in=%d stride=%lld must be multiples of 4", s, g.c_in,
                   (long long)g.src_stride);
        TL_REQUIRE(g.index || g.n_off == 1, "tl_conv_fwd: identity segment %d must have n_off==1", s);
        TL_REQUIRE(g.n_off >= 1 && g.n_off <= 27, "tl_conv_fwd: segment %d: n_off=%d", s, g.n_off);
    }
    TL_REQUIRE((!d.out_act1 || (d.scale1 && d.shift1)) && (!d.out_act2 || (d.scale2 && d.shift2)),
               "tl_conv_fwd: activation output without scale/shift");
    if (mode == TL_MODE_FP32) return conv_fwd_simt(d, stream);
    if (mode == TL_MODE_TF32) return conv_fwd_tc(d, stream, false);
    const bool in4 = d.n_seg == 1 && d.seg[0].c_in == 4 && d.c_out == 32 && d.seg[0].index && d.seg[0].src_stride == 4;
    if (mode == TL_MODE_F16) {
        // A operand in tensor memory (tl_conv_ts.cu) unless TL_TS=0 asks for the shared-memory form of round 1
        // TL_TS: 1 (default) group kernel, cp.async gather + shared-memory-A MMA (tl_conv_grp.cu); 2 tensor-memory-A kernel
        // (tl_conv_ts.cu); 0 round 1's kernel (natural channel order)
        static const int use_ts = getenv("TL_TS") ? atoi(getenv("TL_TS")) : 1;
        if (in4) return conv_fwd_in4(d, stream, 2, use_ts != 0);      // the 4-channel network input is always fp32
        if (use_ts == 1 && conv_halo_eligible(d, (uint32_t)d.src_fp32_mask)) {      // submanifold conv with the level's halo lists
            const int rc = conv_fwd_halo(d, stream, 1);
            if (rc != TL_ERR_UNSUPPORTED) return rc;
        }
        if (use_ts == 1 && conv_grp_eligible(d, (uint32_t)d.src_fp32_mask)) return conv_fwd_grp(d, stream, 1);
        if (use_ts && conv_ts_eligible(d)) return conv_fwd_ts(d, stream, 1, (uint32_t)d.src_fp32_mask);
        TL_REQUIRE(d.src_fp32_mask == 0, "tl_conv_fwd(f16): fp32 segment sources need the tensor-memory path");
        return conv_fwd_tc(d, stream, true);
    }
    if (mode == TL_MODE_F16X2) {
        static const int use_ts2 = getenv("TL_TS") ? atoi(getenv("TL_TS")) : 1;
        if (in4) return conv_fwd_in4(d, stream, 22, true);
        if (use_ts2 == 1 && conv_halo_eligible(d, (uint32_t)d.src_fp32_mask)) {
            const int rc = conv_fwd_halo(d, stream, 2);
            if (rc != TL_ERR_UNSUPPORTED) return rc;
        }
        if (use_ts2 == 1 && conv_grp_eligible(d, (uint32_t)d.src_fp32_mask)) return conv_fwd_grp(d, stream, 2);
        if (!conv_ts_eligible(d)) {
            set_error("tl_conv_fwd(f16x2): shape not eligible (c_in %% 32, c_out %% 32, <= 256 channels)");
            return TL_ERR_UNSUPPORTED;
        }
        return conv_fwd_ts(d, stream, 2, (uint32_t)d.src_fp32_mask);
    }
    set_error("tl_conv_fwd: unknown mode %d", mode);
    return TL_ERR_ARG;
}

int tl_heads_fwd(const void* voxel_feats, int32_t feats_half, const int64_t* v2p, int64_t n, int32_t channels, const float* sem_w1,
                 const float* sem_b1, const float* sem_w2, const float* sem_b2, const float* off_w1,
                 const float* off_b1, const float* off_w2, const float* off_b2, float* backbone_feats,
                 float* sem_logits, float* offsets, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n == 0) return TL_OK;
    TL_REQUIRE(feats_half >= 0 && feats_half <= 3 && (feats_half < 2 || channels % 32 == 0),
               "tl_heads_fwd: feats format %d with %d channels", feats_half, channels);
    const unsigned grid = (unsigned)((n + 127) / 128);
#define TL_HEADS(C)                                                                                                 \
    if (feats_half == 2 && C % 32 == 0)                                                                             \
        k_heads<C, (C % 32 == 0 ? 2 : 0)><<<grid, 128, 0, stream>>>(voxel_feats, v2p, n, sem_w1, sem_b1, sem_w2, sem_b2, off_w1, \
                                                 off_b1, off_w2, off_b2, backbone_feats, sem_logits, offsets);      \
    else if (feats_half == 3 && C % 32 == 0)                                                                        \
        k_heads<C, (C % 32 == 0 ? 3 : 0)><<<grid, 128, 0, stream>>>(voxel_feats, v2p, n, sem_w1, sem_b1, sem_w2, sem_b2, off_w1, \
                                                 off_b1, off_w2, off_b2, backbone_feats, sem_logits, offsets);      \
    else if (feats_half == 1)                                                                                       \
        k_heads<C, 1><<<grid, 128, 0, stream>>>(voxel_feats, v2p, n, sem_w1, sem_b1, sem_w2, sem_b2, off_w1,        \
                                                 off_b1, off_w2, off_b2, backbone_feats, sem_logits, offsets);      \
    else                                                                                                            \
        k_heads<C, 0><<<grid, 128, 0, stream>>>(voxel_feats, v2p, n, sem_w1, sem_b1, sem_w2, sem_b2, off_w1,        \
                                                 off_b1, off_w2, off_b2, backbone_feats, sem_logits, offsets)
    switch (channels) {
        case 8: TL_HEADS(8); break;
        case 16: TL_HEADS(16); break;
        case 32: TL_HEADS(32); break;
        case 64: TL_HEADS(64); break;
        default:
            set_error("tl_heads_fwd: channels=%d unsupported (8/16/32/64)", channels);
            return TL_ERR_UNSUPPORTED;
    }
#undef TL_HEADS
    TL_LAUNCH_CHECK();
    return TL_OK;
}

}  // extern "C"
