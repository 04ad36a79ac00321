// Training-only kernels (SURVEY §8 a12 train mode, a16): BatchNorm batch statistics, BN+ReLU forward and
// backward, and the sparse-conv weight gradient.  The data gradient of every sparse conv reuses tl_conv_fwd
// (same rulebook, transposed weights; offsets mirrored for the 3^3 submanifold table) -- see autograd.py.
//
// Reference semantics: torch.nn.BatchNorm1d(eps=1e-4, momentum=0.1) applied to SparseConvTensor.features
// (tree_learn/model/tree_learn.py:34, blocks.py:57-70): normalise with the biased batch variance, update the
// running variance with the unbiased one; autograd backward (tools/training/train.py:40).
#include <cuda_fp16.h>

#include "tl_common.cuh"

namespace tl {
namespace train {

constexpr int kMaxCB = 16;  // column blocks of 32: channels <= 512 (the widest BatchNorm sees the 2 x 192 skip concat)

// per-channel sum and sum of squares of x [n, c] -> acc[0..c) += sum, acc[c..2c) += sumsq (fp64 atomics)
__global__ void __launch_bounds__(256) k_bn_stats(const float* __restrict__ x, int64_t n, int c, int rows_per_block,
                                                  double* __restrict__ acc) {
    __shared__ float red[2][8][32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r1 = min(r0 + (int64_t)rows_per_block, n);
    for (int cb = 0; cb * 32 < c; ++cb) {
        const int col = cb * 32 + tx;
        float s = 0.f, q = 0.f;
        if (col < c)
            for (int64_t r = r0 + ty; r < r1; r += 8) {
                const float v = __ldg(x + r * c + col);
                s += v;
                q = fmaf(v, v, q);
            }
        red[0][ty][tx] = s;
        red[1][ty][tx] = q;
        __syncthreads();
        if (ty == 0 && col < c) {
            double ds = 0.0, dq = 0.0;
#pragma unroll
            for (int j = 0; j < 8; ++j) ds += (double)red[0][j][tx], dq += (double)red[1][j][tx];
            atomicAdd(acc + col, ds);
            atomicAdd(acc + c + col, dq);
        }
        __syncthreads();
    }
}

// acc -> mean, invstd, scale = gamma*invstd, shift = beta - mean*scale; running stats momentum update
__global__ void k_bn_finalize(const double* __restrict__ acc, int64_t n, int c, const float* __restrict__ gamma,
                              const float* __restrict__ beta, float eps, float momentum, float* running_mean,
                              float* running_var, float* __restrict__ mean, float* __restrict__ invstd,
                              float* __restrict__ scale, float* __restrict__ shift) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= c) return;
    const double m = acc[j] / (double)n;
    double var = acc[c + j] / (double)n - m * m;   // biased (normalisation)
    if (var < 0.0) var = 0.0;
    const float is = (float)(1.0 / sqrt(var + (double)eps));
    const float g = gamma ? gamma[j] : 1.f, b = beta ? beta[j] : 0.f;
    mean[j] = (float)m;
    invstd[j] = is;
    scale[j] = g * is;
    shift[j] = b - (float)m * g * is;
    if (running_mean) running_mean[j] = (1.f - momentum) * running_mean[j] + momentum * (float)m;
    if (running_var) {
        const double unbiased = n > 1 ? var * (double)n / (double)(n - 1) : var;
        running_var[j] = (1.f - momentum) * running_var[j] + momentum * (float)unbiased;
    }
}

// out = relu(scale*x + shift)
__global__ void __launch_bounds__(256) k_bn_relu_apply(const float* __restrict__ x, int64_t total, int c,
                                                       const float* __restrict__ scale, const float* __restrict__ shift,
                                                       float* __restrict__ out) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int col = (int)(e % c);
        out[e] = fmaxf(fmaf(__ldg(x + e), __ldg(scale + col), __ldg(shift + col)), 0.f);
    }
}

// dy = d_act * [scale*x+shift > 0]; acc[0..c) += sum dy, acc[c..2c) += sum dy * xhat
__global__ void __launch_bounds__(256) k_bn_relu_bwd_reduce(const float* __restrict__ x, const float* __restrict__ d_act,
                                                            int64_t n, int c, int rows_per_block,
                                                            const float* __restrict__ scale, const float* __restrict__ shift,
                                                            const float* __restrict__ mean, const float* __restrict__ invstd,
                                                            double* __restrict__ acc) {
    __shared__ float red[2][8][32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r1 = min(r0 + (int64_t)rows_per_block, n);
    for (int cb = 0; cb * 32 < c; ++cb) {
        const int col = cb * 32 + tx;
        float s = 0.f, q = 0.f;
        if (col < c) {
            const float sc = scale[col], sh = shift[col], m = mean[col], is = invstd[col];
            for (int64_t r = r0 + ty; r < r1; r += 8) {
                const float v = __ldg(x + r * c + col);
                const float dy = fmaf(v, sc, sh) > 0.f ? __ldg(d_act + r * c + col) : 0.f;
                s += dy;
                q = fmaf(dy, (v - m) * is, q);
            }
        }
        red[0][ty][tx] = s;
        red[1][ty][tx] = q;
        __syncthreads();
        if (ty == 0 && col < c) {
            double ds = 0.0, dq = 0.0;
#pragma unroll
            for (int j = 0; j < 8; ++j) ds += (double)red[0][j][tx], dq += (double)red[1][j][tx];
            atomicAdd(acc + col, ds);
            atomicAdd(acc + c + col, dq);
        }
        __syncthreads();
    }
}

// batch-stat mode: dx = scale * (dy - sum_dy/n - xhat * sum_dy_xhat/n); frozen (eval) mode: dx = scale * dy
__global__ void __launch_bounds__(256) k_bn_relu_bwd_apply(const float* __restrict__ x, const float* __restrict__ d_act,
                                                           int64_t n, int c, const float* __restrict__ scale,
                                                           const float* __restrict__ shift, const float* __restrict__ mean,
                                                           const float* __restrict__ invstd, const double* __restrict__ acc,
                                                           int batch_stats, float* __restrict__ dx,
                                                           float* __restrict__ dgamma, float* __restrict__ dbeta) {
    const int64_t total = n * c;
    const double inv_n = 1.0 / (double)n;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int col = (int)(e % c);
        const float v = __ldg(x + e);
        const float sc = __ldg(scale + col);
        const float dy = fmaf(v, sc, __ldg(shift + col)) > 0.f ? __ldg(d_act + e) : 0.f;
        float g = dy;
        if (batch_stats) {
            const float xhat = (v - __ldg(mean + col)) * __ldg(invstd + col);
            g = dy - (float)(acc[col] * inv_n) - xhat * (float)(acc[c + col] * inv_n);
        }
        dx[e] = sc * g;
    }
    if (blockIdx.x == 0)
        for (int j = threadIdx.x; j < c; j += blockDim.x) {
            if (dbeta) dbeta[j] = (float)acc[j];
            if (dgamma) dgamma[j] = (float)acc[c + j];
        }
}

// ---- float4 forms (c % 4 == 0, c <= 512; rows are 16 B aligned): the scalar kernels above drew 0.46-0.75 TB/s of HBM
// (profiles/r02_ncu_launches_summary_train.txt: one 4 B load per lane and iteration, an int64 modulo per element)
// One thread keeps ONE float4 column (t % c4) and strides over the rows: S = the largest multiple of c4 <= 256 threads
// are active, S / c4 rows per pass; fp32 partial sums per thread, fp64 across threads and blocks.
template <bool BWD>
__global__ void __launch_bounds__(256) k_bn_reduce_v4(const float* __restrict__ x, const float* __restrict__ d_act, int64_t n, int c,
                                                      int rows_per_block, const float* __restrict__ scale,
                                                      const float* __restrict__ shift, const float* __restrict__ mean,
                                                      const float* __restrict__ invstd, double* __restrict__ acc) {
    __shared__ float4 red[2][256];
    const int c4 = c >> 2, t = threadIdx.x;
    const int rpp = 256 / c4, S = rpp * c4;
    const int col4 = t % c4, roff = t / c4;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r1 = min(r0 + (int64_t)rows_per_block, n);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
    if (t < S) {
        const float4* x4 = reinterpret_cast<const float4*>(x) + col4;
        const float4* d4 = reinterpret_cast<const float4*>(d_act) + col4;
        float4 sc = s, sh = s, m = s, is = s;
        if (BWD) {
            sc = __ldg(reinterpret_cast<const float4*>(scale) + col4), sh = __ldg(reinterpret_cast<const float4*>(shift) + col4);
            m = __ldg(reinterpret_cast<const float4*>(mean) + col4), is = __ldg(reinterpret_cast<const float4*>(invstd) + col4);
        }
#pragma unroll 4
        for (int64_t r = r0 + roff; r < r1; r += rpp) {
            const float4 v = __ldg(x4 + r * c4);
            if (BWD) {
                const float4 g = __ldg(d4 + r * c4);
                const float dx_ = fmaf(v.x, sc.x, sh.x) > 0.f ? g.x : 0.f, dy_ = fmaf(v.y, sc.y, sh.y) > 0.f ? g.y : 0.f;
                const float dz_ = fmaf(v.z, sc.z, sh.z) > 0.f ? g.z : 0.f, dw_ = fmaf(v.w, sc.w, sh.w) > 0.f ? g.w : 0.f;
                s.x += dx_, s.y += dy_, s.z += dz_, s.w += dw_;
                q.x = fmaf(dx_, (v.x - m.x) * is.x, q.x), q.y = fmaf(dy_, (v.y - m.y) * is.y, q.y);
                q.z = fmaf(dz_, (v.z - m.z) * is.z, q.z), q.w = fmaf(dw_, (v.w - m.w) * is.w, q.w);
            } else {
                s.x += v.x, s.y += v.y, s.z += v.z, s.w += v.w;
                q.x = fmaf(v.x, v.x, q.x), q.y = fmaf(v.y, v.y, q.y), q.z = fmaf(v.z, v.z, q.z), q.w = fmaf(v.w, v.w, q.w);
            }
        }
    }
    red[0][t] = s, red[1][t] = q;
    __syncthreads();
    if (t < c4) {
        double ds[4] = {0, 0, 0, 0}, dq[4] = {0, 0, 0, 0};
        for (int j = 0; j < rpp; ++j) {
            const float4 a = red[0][j * c4 + t], b = red[1][j * c4 + t];
            ds[0] += (double)a.x, ds[1] += (double)a.y, ds[2] += (double)a.z, ds[3] += (double)a.w;
            dq[0] += (double)b.x, dq[1] += (double)b.y, dq[2] += (double)b.z, dq[3] += (double)b.w;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            atomicAdd(acc + 4 * t + i, ds[i]);
            atomicAdd(acc + c + 4 * t + i, dq[i]);
        }
    }
}

// elementwise passes over [n, c] as float4 columns: MODE 0 out = relu(scale x + shift); MODE 1 the BN+ReLU input gradient
template <int MODE>
__global__ void __launch_bounds__(256) k_bn_elem_v4(const float* __restrict__ x, const float* __restrict__ d_act, int64_t n, int c,
                                                    const float* __restrict__ scale, const float* __restrict__ shift,
                                                    const float* __restrict__ mean, const float* __restrict__ invstd,
                                                    const double* __restrict__ acc, int batch_stats, float* __restrict__ out,
                                                    float* __restrict__ dgamma, float* __restrict__ dbeta) {
    const int c4 = c >> 2, t = threadIdx.x;
    const int rpp = 256 / c4, S = rpp * c4;
    if (MODE == 1 && blockIdx.x == 0)
        for (int j = t; j < c; j += 256) {
            if (dbeta) dbeta[j] = (float)acc[j];
            if (dgamma) dgamma[j] = (float)acc[c + j];
        }
    if (t >= S) return;
    const int col4 = t % c4, roff = t / c4;
    const float4 sc = __ldg(reinterpret_cast<const float4*>(scale) + col4), sh = __ldg(reinterpret_cast<const float4*>(shift) + col4);
    float4 m = sc, is = sc, a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
    if (MODE == 1 && batch_stats) {
        m = __ldg(reinterpret_cast<const float4*>(mean) + col4), is = __ldg(reinterpret_cast<const float4*>(invstd) + col4);
        const double inv_n = 1.0 / (double)n;
        a0 = make_float4((float)(acc[4 * col4] * inv_n), (float)(acc[4 * col4 + 1] * inv_n), (float)(acc[4 * col4 + 2] * inv_n),
                         (float)(acc[4 * col4 + 3] * inv_n));
        a1 = make_float4((float)(acc[c + 4 * col4] * inv_n), (float)(acc[c + 4 * col4 + 1] * inv_n), (float)(acc[c + 4 * col4 + 2] * inv_n),
                         (float)(acc[c + 4 * col4 + 3] * inv_n));
    }
    const float4* x4 = reinterpret_cast<const float4*>(x) + col4;
    const float4* d4 = reinterpret_cast<const float4*>(d_act) + col4;
    float4* o4 = reinterpret_cast<float4*>(out) + col4;
#pragma unroll 4
    for (int64_t r = (int64_t)blockIdx.x * rpp + roff; r < n; r += (int64_t)gridDim.x * rpp) {
        const float4 v = __ldg(x4 + r * c4);
        float4 o;
        if (MODE == 0) {
            o = make_float4(fmaxf(fmaf(v.x, sc.x, sh.x), 0.f), fmaxf(fmaf(v.y, sc.y, sh.y), 0.f), fmaxf(fmaf(v.z, sc.z, sh.z), 0.f),
                            fmaxf(fmaf(v.w, sc.w, sh.w), 0.f));
        } else {
            const float4 g = __ldg(d4 + r * c4);
            float4 dy = make_float4(fmaf(v.x, sc.x, sh.x) > 0.f ? g.x : 0.f, fmaf(v.y, sc.y, sh.y) > 0.f ? g.y : 0.f,
                                    fmaf(v.z, sc.z, sh.z) > 0.f ? g.z : 0.f, fmaf(v.w, sc.w, sh.w) > 0.f ? g.w : 0.f);
            if (batch_stats) {
                dy.x = dy.x - a0.x - (v.x - m.x) * is.x * a1.x, dy.y = dy.y - a0.y - (v.y - m.y) * is.y * a1.y;
                dy.z = dy.z - a0.z - (v.z - m.z) * is.z * a1.z, dy.w = dy.w - a0.w - (v.w - m.w) * is.w * a1.w;
            }
            o = make_float4(sc.x * dy.x, sc.y * dy.y, sc.z * dy.z, sc.w * dy.w);
        }
        o4[r * c4] = o;
    }
}

// ------------------------------------------------------------------------------------------------
// weight gradient: dw[k][ci][co] += sum_r src[index[k][r], ci] * d_out[r, co]
// One CTA (64 threads, 4x4 register tile each) owns a 32x32 (ci, co) tile of one offset k over a slab of rows;
// gathered src rows and d_out rows are staged 32 rows at a time in shared memory; partial sums -> fp32 red.add.
// ------------------------------------------------------------------------------------------------
constexpr int WG_ROWS = 32;

__global__ void __launch_bounds__(64) k_conv_wgrad(const float* __restrict__ src, int64_t src_stride, int c_in, int n_off,
                                                   const int32_t* __restrict__ index, int64_t index_stride,
                                                   const uint32_t* __restrict__ tile_mask, const float* __restrict__ d_out,
                                                   int64_t n_out, int c_out, int rows_per_block, int co_tiles,
                                                   float* __restrict__ dw) {
    __shared__ __align__(16) float xs[WG_ROWS][32];
    __shared__ __align__(16) float ys[WG_ROWS][32];
    const int tid = threadIdx.x;
    const int k = blockIdx.y;
    const int ci0 = (blockIdx.z / co_tiles) * 32, co0 = (blockIdx.z % co_tiles) * 32;
    const int ti = tid >> 3, tj = tid & 7;
    const int64_t r_begin = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r_end = min(r_begin + (int64_t)rows_per_block, n_out);
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    const int lrow = tid >> 3, lc = (tid & 7) * 4;   // loader: 8 rows x 8 float4 per pass, 4 passes
    // software pipeline: the rows of batch i+1 are fetched into registers while batch i is multiplied out of shared memory
    float4 xr[4], yr[4];
    auto live = [&](int64_t rb) {   // block-uniform: does the 128-row tile of this batch have offset k at all?
        return !(index && tile_mask) || ((tile_mask[rb / TL_TILE_ROWS] >> k) & 1u);
    };
    auto fetch = [&](int64_t rb) {
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int64_t r = rb + p * 8 + lrow;
            float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), yv = xv;
            if (r < r_end) {
                const int64_t s = index ? (int64_t)__ldg(index + (int64_t)k * index_stride + r) : r;
                if (s >= 0) {
                    const float* xp = src + s * src_stride + ci0 + lc;
                    const float* yp = d_out + r * c_out + co0 + lc;
                    if (ci0 + lc + 3 < c_in) xv = __ldg(reinterpret_cast<const float4*>(xp));
                    else {
                        if (ci0 + lc < c_in) xv.x = __ldg(xp);
                        if (ci0 + lc + 1 < c_in) xv.y = __ldg(xp + 1);
                        if (ci0 + lc + 2 < c_in) xv.z = __ldg(xp + 2);
                    }
                    if (co0 + lc + 3 < c_out) yv = __ldg(reinterpret_cast<const float4*>(yp));
                    else {
                        if (co0 + lc < c_out) yv.x = __ldg(yp);
                        if (co0 + lc + 1 < c_out) yv.y = __ldg(yp + 1);
                        if (co0 + lc + 2 < c_out) yv.z = __ldg(yp + 2);
                    }
                }
            }
            xr[p] = xv, yr[p] = yv;
        }
    };
    auto next_live = [&](int64_t rb) {
        while (rb < r_end && !live(rb)) rb += WG_ROWS;
        return rb;
    };
    int64_t rb = next_live(r_begin);
    if (rb < r_end) fetch(rb);
    while (rb < r_end) {
        __syncthreads();
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            *reinterpret_cast<float4*>(&xs[p * 8 + lrow][lc]) = xr[p];
            *reinterpret_cast<float4*>(&ys[p * 8 + lrow][lc]) = yr[p];
        }
        __syncthreads();
        const int64_t nb = next_live(rb + WG_ROWS);
        if (nb < r_end) fetch(nb);
#pragma unroll 8
        for (int rr = 0; rr < WG_ROWS; ++rr) {
            const float4 a = *reinterpret_cast<const float4*>(&xs[rr][ti * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&ys[rr][tj * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(av[u], bv[v], acc[u][v]);
        }
        rb = nb;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int ci = ci0 + ti * 4 + u;
        if (ci >= c_in) continue;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const int co = co0 + tj * 4 + v;
            if (co < c_out && acc[u][v] != 0.f) atomicAdd(dw + ((int64_t)k * c_in + ci) * c_out + co, acc[u][v]);
        }
    }
}


// TF32 tensor-core variant (mma.sync.m16n8k8, fp32 accumulate) of the weight gradient, used by the tf32 / f16 training
// modes: same tiling and software pipeline as k_conv_wgrad; warp w owns 16 of the 32 input channels (M), all 32 output
// channels (4 N tiles of 8), K = the 32 staged rows in 4 steps of 8.  Operands are rounded to TF32 (RN) when staged.
// Rows are padded to 40 floats so that the fragment loads (lane = 4*group + tig -> row tig, column group) are conflict-free.
constexpr int WG_PAD = 40;
__device__ __forceinline__ float to_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}
__global__ void __launch_bounds__(64) k_conv_wgrad_tf32(const float* __restrict__ src, int64_t src_stride, int c_in, int n_off,
                                                        const int32_t* __restrict__ index, int64_t index_stride,
                                                        const uint32_t* __restrict__ tile_mask,
                                                        const float* __restrict__ d_out, int64_t n_out, int c_out,
                                                        int rows_per_block, int co_tiles, float* __restrict__ dw) {
    __shared__ __align__(16) float xs[WG_ROWS][WG_PAD];
    __shared__ __align__(16) float ys[WG_ROWS][WG_PAD];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, grp = lane >> 2, tig = lane & 3;
    const int k = blockIdx.y;
    const int ci0 = (blockIdx.z / co_tiles) * 32, co0 = (blockIdx.z % co_tiles) * 32;
    const int64_t r_begin = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r_end = min(r_begin + (int64_t)rows_per_block, n_out);
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    const int lrow = tid >> 3, lc = (tid & 7) * 4;   // loader: 8 rows x 8 float4 per pass, 4 passes
    float4 xr[4], yr[4];
    auto live = [&](int64_t rb) { return !(index && tile_mask) || ((tile_mask[rb / TL_TILE_ROWS] >> k) & 1u); };
    auto fetch = [&](int64_t rb) {
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int64_t r = rb + p * 8 + lrow;
            float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), yv = xv;
            if (r < r_end) {
                const int64_t s = index ? (int64_t)__ldg(index + (int64_t)k * index_stride + r) : r;
                if (s >= 0) {
                    const float* xp = src + s * src_stride + ci0 + lc;
                    const float* yp = d_out + r * c_out + co0 + lc;
                    if (ci0 + lc + 3 < c_in) xv = __ldg(reinterpret_cast<const float4*>(xp));
                    else {
                        if (ci0 + lc < c_in) xv.x = __ldg(xp);
                        if (ci0 + lc + 1 < c_in) xv.y = __ldg(xp + 1);
                        if (ci0 + lc + 2 < c_in) xv.z = __ldg(xp + 2);
                    }
                    if (co0 + lc + 3 < c_out) yv = __ldg(reinterpret_cast<const float4*>(yp));
                    else {
                        if (co0 + lc < c_out) yv.x = __ldg(yp);
                        if (co0 + lc + 1 < c_out) yv.y = __ldg(yp + 1);
                        if (co0 + lc + 2 < c_out) yv.z = __ldg(yp + 2);
                    }
                }
            }
            xr[p] = xv, yr[p] = yv;
        }
    };
    auto next_live = [&](int64_t rb) {
        while (rb < r_end && !live(rb)) rb += WG_ROWS;
        return rb;
    };
    int64_t rb = next_live(r_begin);
    if (rb < r_end) fetch(rb);
    while (rb < r_end) {
        __syncthreads();
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            *reinterpret_cast<float4*>(&xs[p * 8 + lrow][lc]) =
                make_float4(to_tf32(xr[p].x), to_tf32(xr[p].y), to_tf32(xr[p].z), to_tf32(xr[p].w));
            *reinterpret_cast<float4*>(&ys[p * 8 + lrow][lc]) =
                make_float4(to_tf32(yr[p].x), to_tf32(yr[p].y), to_tf32(yr[p].z), to_tf32(yr[p].w));
        }
        __syncthreads();
        const int64_t nb = next_live(rb + WG_ROWS);
        if (nb < r_end) fetch(nb);
#pragma unroll
        for (int ks = 0; ks < WG_ROWS; ks += 8) {
            // A[m = ci][k = row] = xs[row][ci]: a0 (m = grp, k = tig), a1 (m = grp + 8, k = tig), a2 / a3: k + 4
            const uint32_t a0 = __float_as_uint(xs[ks + tig][warp * 16 + grp]);
            const uint32_t a1 = __float_as_uint(xs[ks + tig][warp * 16 + grp + 8]);
            const uint32_t a2 = __float_as_uint(xs[ks + tig + 4][warp * 16 + grp]);
            const uint32_t a3 = __float_as_uint(xs[ks + tig + 4][warp * 16 + grp + 8]);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                // B[k = row][n = co] = ys[row][co]: b0 (k = tig, n = grp), b1 (k = tig + 4, n = grp)
                const uint32_t b0 = __float_as_uint(ys[ks + tig][nt * 8 + grp]);
                const uint32_t b1 = __float_as_uint(ys[ks + tig + 4][nt * 8 + grp]);
                asm volatile(
                    "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
                    "{%0, %1, %2, %3};"
                    : "+f"(acc[nt][0]), "+f"(acc[nt][1]), "+f"(acc[nt][2]), "+f"(acc[nt][3])
                    : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            }
        }
        rb = nb;
    }
    // D fragment: c0 (m = grp, n = 2 tig), c1 (m = grp, n = 2 tig + 1), c2 / c3: m + 8
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int ci = ci0 + warp * 16 + grp + (e >> 1) * 8, co = co0 + nt * 8 + 2 * tig + (e & 1);
            if (ci < c_in && co < c_out && acc[nt][e] != 0.f) atomicAdd(dw + ((int64_t)k * c_in + ci) * c_out + co, acc[nt][e]);
        }
}

// ------------------------------------------------------------------------------------------------
// weight packing for the tcgen05 paths (training repacks every conv twice per step: forward and data-gradient layout)
//   src: the parameter, spconv KRSC layout [C_out][K][C_in] fp32
//   dst: [K][rows_k / ... ] B-operand image [K][c_in'/bk][c_out'][bk] with the UMMA swizzle applied per row
//        (sparse.pack_weight_tc), where for the forward conv (c_out', c_in') = (C_out, C_in) and element
//        (k, n, c) = W[n][k][c]; for the data gradient (transpose = 1) (c_out', c_in') = (C_in, C_out) and
//        (k, n, c) = W[c][mirror ? K-1-k : k][n].   fp32 rounded to TF32 (RN-even) or fp16.
// ------------------------------------------------------------------------------------------------
template <bool HALF>
__global__ void __launch_bounds__(256) k_pack_weight_tc(const float* __restrict__ w, int c_out, int n_off, int c_in,
                                                        int transpose, int mirror, int bk, void* __restrict__ out) {
    const int co_p = transpose ? c_in : c_out, ci_p = transpose ? c_out : c_in;     // packed conv's C_out', C_in'
    const int64_t total = (int64_t)n_off * co_p * ci_p;
    const int eb = HALF ? 2 : 4, epc = 16 / eb, ch = bk / epc;                       // elements per 16 B chunk, chunks per row
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        // destination coordinates: [k][kb][n][chunk][elem]
        int64_t t = e;
        const int el = (int)(t % epc);
        t /= epc;
        const int cdst = (int)(t % ch);
        t /= ch;
        const int n = (int)(t % co_p);
        t /= co_p;
        const int kb = (int)(t % (ci_p / bk));
        const int k = (int)(t / (ci_p / bk));
        const int x = (bk * eb == 128) ? (n & 7) : ((n >> 1) & 3);
        const int c = kb * bk + (cdst ^ x) * epc + el;                               // source channel of the packed conv
        const int ks = (transpose && mirror) ? n_off - 1 - k : k;
        const float v = transpose ? w[((int64_t)c * n_off + ks) * c_in + n] : w[((int64_t)n * n_off + ks) * c_in + c];
        if (HALF) {
            reinterpret_cast<__half*>(out)[e] = __float2half_rn(v);
        } else {
            uint32_t i = __float_as_uint(v);
            i = (i + 0x0FFFu + ((i >> 13) & 1u)) & ~0x1FFFu;                         // round to nearest even at 10 mantissa bits
            reinterpret_cast<uint32_t*>(out)[e] = i;
        }
    }
}

static inline int rows_per_block_for(int64_t n, int target_blocks) {
    int64_t rpb = (n + target_blocks - 1) / target_blocks;
    rpb = (rpb + 127) / 128 * 128;
    return (int)(rpb < 128 ? 128 : rpb);
}

}  // namespace train
}  // namespace tl

using namespace tl;

extern "C" {

int tl_bn_stats(const float* x, int64_t n, int32_t c, double* acc, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    TL_REQUIRE(x && acc && n > 0 && c > 0 && c <= 32 * train::kMaxCB, "tl_bn_stats: n=%lld c=%d", (long long)n, c);
    TL_CUDA_CHECK(cudaMemsetAsync(acc, 0, sizeof(double) * 2 * c, stream));
    const int rpb = train::rows_per_block_for(n, 148 * 8);
    if (c % 4 == 0)
        train::k_bn_reduce_v4<false><<<(unsigned)((n + rpb - 1) / rpb), 256, 0, stream>>>(x, nullptr, n, c, rpb, nullptr, nullptr, nullptr,
                                                                                        nullptr, acc);
    else
        train::k_bn_stats<<<(unsigned)((n + rpb - 1) / rpb), 256, 0, stream>>>(x, n, c, rpb, acc);
    TL_LAUNCH_CHECK();
    return TL_OK;
}

int tl_bn_finalize(const double* acc, int64_t n, int32_t c, const float* gamma, const float* beta, float eps,
                   float momentum, float* running_mean, float* running_var, float* mean, float* invstd, float* scale,
                   float* shift, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    TL_REQUIRE(acc && mean && invstd && scale && shift && n > 0 && c > 0, "tl_bn_finalize: bad arguments");
    train::k_bn_finalize<<<(c + 127) / 128, 128, 0, stream>>>(acc, n, c, gamma, beta, eps, momentum, running_mean,
                                                              running_var, mean, invstd, scale, shift);
    TL_LAUNCH_CHECK();
    return TL_OK;
}

int tl_bn_relu_apply(const float* x, int64_t n, int32_t c, const float* scale, const float* shift, float* out,
                     void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n == 0) return TL_OK;
    TL_REQUIRE(x && scale && shift && out && c > 0, "tl_bn_relu_apply: bad arguments");
    const int64_t total = n * c;
    const unsigned grid = (unsigned)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
    if (c % 4 == 0 && c <= 512) {
        const int64_t passes = (n + (256 / (c / 4)) - 1) / (256 / (c / 4));
        train::k_bn_elem_v4<0><<<(unsigned)(passes > 148 * 16 ? 148 * 16 : passes), 256, 0, stream>>>(
            x, nullptr, n, c, scale, shift, nullptr, nullptr, nullptr, 0, out, nullptr, nullptr);
    } else {
        train::k_bn_relu_apply<<<grid, 256, 0, stream>>>(x, total, c, scale, shift, out);
    }
    TL_LAUNCH_CHECK();
    return TL_OK;
}

int tl_bn_relu_bwd(const float* x, const float* d_act, int64_t n, int32_t c, const float* scale, const float* shift,
                   const float* mean, const float* invstd, int32_t batch_stats, double* acc, float* dx, float* dgamma,
                   float* dbeta, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n == 0) return TL_OK;
    TL_REQUIRE(x && d_act && scale && shift && mean && invstd && acc && dx && c > 0 && c <= 32 * train::kMaxCB,
               "tl_bn_relu_bwd: bad arguments");
    TL_CUDA_CHECK(cudaMemsetAsync(acc, 0, sizeof(double) * 2 * c, stream));
    const int rpb = train::rows_per_block_for(n, 148 * 8);
    if (c % 4 == 0)
        train::k_bn_reduce_v4<true><<<(unsigned)((n + rpb - 1) / rpb), 256, 0, stream>>>(x, d_act, n, c, rpb, scale, shift, mean,
                                                                                       invstd, acc);
    else
        train::k_bn_relu_bwd_reduce<<<(unsigned)((n + rpb - 1) / rpb), 256, 0, stream>>>(x, d_act, n, c, rpb, scale, shift,
                                                                                       mean, invstd, acc);
    TL_LAUNCH_CHECK();
    const int64_t total = n * c;
    const unsigned grid = (unsigned)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
    if (c % 4 == 0) {
        const int64_t passes = (n + (256 / (c / 4)) - 1) / (256 / (c / 4));
        train::k_bn_elem_v4<1><<<(unsigned)(passes > 148 * 16 ? 148 * 16 : passes), 256, 0, stream>>>(
            x, d_act, n, c, scale, shift, mean, invstd, acc, batch_stats, dx, dgamma, dbeta);
    } else {
        train::k_bn_relu_bwd_apply<<<grid, 256, 0, stream>>>(x, d_act, n, c, scale, shift, mean, invstd, acc, batch_stats, dx,
                                                             dgamma, dbeta);
    }
    TL_LAUNCH_CHECK();
    return TL_OK;
}

int tl_conv_wgrad(const float* src, int64_t src_stride, int32_t c_in, int32_t n_off, const int32_t* index,
                  int64_t index_stride, const uint32_t* tile_mask, const float* d_out, int64_t n_out, int32_t c_out,
                  float* dw, int32_t tf32, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    TL_REQUIRE(src && d_out && dw && c_in > 0 && c_out > 0 && n_off >= 1 && n_off <= 27, "tl_conv_wgrad: bad arguments");
    TL_REQUIRE(index || n_off == 1, "tl_conv_wgrad: identity map needs n_off == 1");
    TL_REQUIRE(src_stride % 4 == 0 && c_out % 4 == 0, "tl_conv_wgrad: src_stride and c_out must be multiples of 4");
    TL_CUDA_CHECK(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)n_off * c_in * c_out, stream));
    if (n_out == 0) return TL_OK;
    const int ci_tiles = (c_in + 31) / 32, co_tiles = (c_out + 31) / 32;
    int target = 148 * 24 / (n_off * ci_tiles * co_tiles);
    if (target < 1) target = 1;
    const int rpb = train::rows_per_block_for(n_out, target);
    dim3 grid((unsigned)((n_out + rpb - 1) / rpb), (unsigned)n_off, (unsigned)(ci_tiles * co_tiles));
    if (tf32)
        train::k_conv_wgrad_tf32<<<grid, 64, 0, stream>>>(src, src_stride, c_in, n_off, index, index_stride, tile_mask, d_out,
                                                          n_out, c_out, rpb, co_tiles, dw);
    else
        train::k_conv_wgrad<<<grid, 64, 0, stream>>>(src, src_stride, c_in, n_off, index, index_stride, tile_mask, d_out,
                                                     n_out, c_out, rpb, co_tiles, dw);
    TL_LAUNCH_CHECK();
    return TL_OK;
}

int tl_pack_weight_tc(const float* w, int32_t c_out, int32_t n_off, int32_t c_in, int32_t transpose, int32_t mirror,
                      int32_t half, int32_t bk, void* out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    const int co_p = transpose ? c_in : c_out, ci_p = transpose ? c_out : c_in;
    TL_REQUIRE(w && out && n_off >= 1 && (bk == 32 || bk == 64) && ci_p % bk == 0 && co_p % 32 == 0,
               "tl_pack_weight_tc: c_out=%d n_off=%d c_in=%d transpose=%d bk=%d", c_out, n_off, c_in, transpose, bk);
    const int64_t total = (int64_t)n_off * c_out * c_in;
    const unsigned grid = (unsigned)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
    if (half) train::k_pack_weight_tc<true><<<grid, 256, 0, stream>>>(w, c_out, n_off, c_in, transpose, mirror, bk, out);
    else train::k_pack_weight_tc<false><<<grid, 256, 0, stream>>>(w, c_out, n_off, c_in, transpose, mirror, bk, out);
    TL_LAUNCH_CHECK();
    return TL_OK;
}

}  // extern "C"
