// Post-processing joins that follow the per-tile path: exact-coordinate hash join for the
// voxel -> voxel prediction propagation, and the (prediction x ground-truth) co-occurrence counts behind the
// instance IoU matrix.  HBM-bound integer work: one pass over the rows, atomics into small tables.
#include "tl_common.cuh"

namespace {

// ---- coordinate keys ------------------------------------------------------------------------------------------------
// A key is the row's three coordinates as fp64 values, optionally rounded to two decimals the way numpy.round does it
// (multiply by 100, rint, divide by 100, all in the array's own dtype; tree_learn/util/pipeline.py:443,458), with -0.0
// folded onto +0.0 (Python's tuple hash / equality treats them as the same key).
struct Key3 {
    unsigned long long a, b, c;
};

__device__ __forceinline__ double round2_f64(double v) { return __ddiv_rn(rint(__dmul_rn(v, 100.0)), 100.0); }
__device__ __forceinline__ float round2_f32(float v) { return __fdiv_rn(rintf(__fmul_rn(v, 100.0f)), 100.0f); }

__device__ __forceinline__ unsigned long long canon(double v) {
    if (v == 0.0) v = 0.0;
    return (unsigned long long)__double_as_longlong(v);
}

__device__ __forceinline__ Key3 load_key(const void* xyz, int is_f64, int round2, int64_t i) {
    double v[3];
    if (is_f64) {
        const double* p = (const double*)xyz + 3 * i;
#pragma unroll
        for (int d = 0; d < 3; ++d) v[d] = round2 ? round2_f64(p[d]) : p[d];
    } else {
        const float* p = (const float*)xyz + 3 * i;
#pragma unroll
        for (int d = 0; d < 3; ++d) v[d] = (double)(round2 ? round2_f32(p[d]) : p[d]);
    }
    return Key3{canon(v[0]), canon(v[1]), canon(v[2])};
}

__device__ __forceinline__ bool same(const Key3& x, const Key3& y) { return x.a == y.a && x.b == y.b && x.c == y.c; }

__device__ __forceinline__ unsigned long long mix64(unsigned long long h) {
    h ^= h >> 33;
    h *= 0xff51afd7ed558ccdull;
    h ^= h >> 33;
    h *= 0xc4ceb9fe1a85ec53ull;
    h ^= h >> 33;
    return h;
}
__device__ __forceinline__ unsigned long long slot_of(const Key3& k, unsigned long long mask) {
    return mix64(k.a ^ mix64(k.b ^ mix64(k.c + 0x9e3779b97f4a7c15ull))) & mask;
}

__global__ void k_fill_i64(long long* p, int64_t n, long long v) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

// Open-addressing table: rep[s] = first build row that claimed slot s (its coordinates ARE the slot's key),
// last[s] = highest build row with that key (a Python dict built by zip keeps the LAST value of a repeated key).
__global__ void k_join_build(const void* __restrict__ xyz, int is_f64, int round2, int64_t n, long long* rep,
                             long long* last, unsigned long long mask) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Key3 key = load_key(xyz, is_f64, round2, i);
    unsigned long long s = slot_of(key, mask);
    for (unsigned long long probe = 0; probe <= mask; ++probe, s = (s + 1) & mask) {
        long long r = atomicCAS((unsigned long long*)&rep[s], (unsigned long long)-1ll, (unsigned long long)i);
        if (r == -1ll || same(load_key(xyz, is_f64, round2, r), key)) {
            atomicMax(&last[s], (long long)i);
            return;
        }
    }
}

__global__ void k_join_probe(const void* __restrict__ build_xyz, int build_f64, int build_round2,
                             const long long* __restrict__ rep, const long long* __restrict__ last,
                             unsigned long long mask, const int64_t* __restrict__ build_vals,
                             const void* __restrict__ probe_xyz, int probe_f64, int probe_round2, int64_t n_probe,
                             int64_t missing, int64_t* __restrict__ out) {
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= n_probe) return;
    const Key3 key = load_key(probe_xyz, probe_f64, probe_round2, j);
    unsigned long long s = slot_of(key, mask);
    int64_t res = missing;
    for (unsigned long long probe = 0; probe <= mask; ++probe, s = (s + 1) & mask) {
        const long long r = rep[s];
        if (r == -1ll) break;
        if (same(load_key(build_xyz, build_f64, build_round2, r), key)) {
            res = build_vals[last[s]];
            break;
        }
    }
    out[j] = res;
}

// ---- co-occurrence counts ----------------------------------------------------------------------------------------
// counts[(p', g')] += 1 per point, p' = pred if 0 <= pred < n_pred else n_pred, g' likewise: the extra row / column
// keep the points whose other label is out of range, so row and column sums are the full mask sizes.
__global__ void k_cooccurrence(const int64_t* __restrict__ pred, const int64_t* __restrict__ gt, int64_t n,
                               int64_t n_pred, int64_t n_gt, unsigned long long* __restrict__ counts) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const bool live = i < n;
    long long cell = -1;
    if (live) {
        int64_t p = pred[i], g = gt[i];
        if (p < 0 || p >= n_pred) p = n_pred;
        if (g < 0 || g >= n_gt) g = n_gt;
        cell = p * (n_gt + 1) + g;
    }
    // neighbouring points mostly share both labels: one atomic per distinct cell in the warp
    const unsigned peers = __match_any_sync(0xffffffffu, cell);
    if (live && (int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&counts[cell], (unsigned long long)__popc(peers));
}

}  // namespace

extern "C" {

size_t tl_hash_join_workspace_bytes(int64_t n_build) {
    unsigned long long cap = 64;
    while (cap < 2ull * (unsigned long long)(n_build > 0 ? n_build : 0)) cap <<= 1;
    return 2 * tl::align_up(cap * sizeof(long long)) + 256;
}

int tl_hash_join_last(const void* build_xyz, int32_t build_f64, int32_t build_round2, const int64_t* build_vals,
                      int64_t n_build, const void* probe_xyz, int32_t probe_f64, int32_t probe_round2, int64_t n_probe,
                      int64_t missing, int64_t* out_vals, void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    TL_REQUIRE(n_build >= 0 && n_probe >= 0, "tl_hash_join_last: negative row count");
    if (n_probe == 0) return TL_OK;
    TL_REQUIRE(probe_xyz && out_vals, "tl_hash_join_last: null probe/out pointer");
    TL_REQUIRE(n_build == 0 || (build_xyz && build_vals), "tl_hash_join_last: null build pointer");
    TL_REQUIRE(workspace && workspace_bytes >= tl_hash_join_workspace_bytes(n_build),
               "tl_hash_join_last: workspace too small (%zu < %zu)", workspace_bytes, tl_hash_join_workspace_bytes(n_build));
    unsigned long long cap = 64;
    while (cap < 2ull * (unsigned long long)n_build) cap <<= 1;
    tl::Carver cv(workspace, workspace_bytes);
    long long* rep = cv.take<long long>(cap);
    long long* last = cv.take<long long>(cap);
    const unsigned fill_blocks = (unsigned)((cap + 255) / 256 < 148 * 8 ? (cap + 255) / 256 : 148 * 8);
    k_fill_i64<<<fill_blocks, 256, 0, stream>>>(rep, (int64_t)cap, -1ll);
    TL_LAUNCH_CHECK();
    k_fill_i64<<<fill_blocks, 256, 0, stream>>>(last, (int64_t)cap, -1ll);
    TL_LAUNCH_CHECK();
    if (n_build > 0) {
        k_join_build<<<(unsigned)((n_build + 255) / 256), 256, 0, stream>>>(build_xyz, build_f64, build_round2, n_build,
                                                                            rep, last, cap - 1);
        TL_LAUNCH_CHECK();
    }
    k_join_probe<<<(unsigned)((n_probe + 255) / 256), 256, 0, stream>>>(build_xyz, build_f64, build_round2, rep, last,
                                                                        cap - 1, build_vals, probe_xyz, probe_f64,
                                                                        probe_round2, n_probe, missing, out_vals);
    TL_LAUNCH_CHECK();
    return TL_OK;
}

int tl_cooccurrence_counts(const int64_t* pred, const int64_t* gt, int64_t n, int64_t n_pred, int64_t n_gt,
                           unsigned long long* counts, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    TL_REQUIRE(n >= 0 && n_pred >= 0 && n_gt >= 0, "tl_cooccurrence_counts: negative size");
    TL_REQUIRE(counts, "tl_cooccurrence_counts: null counts");
    TL_REQUIRE((double)(n_pred + 1) * (double)(n_gt + 1) < 4.0e9, "tl_cooccurrence_counts: table (%lld+1)x(%lld+1) too large",
               (long long)n_pred, (long long)n_gt);
    TL_CUDA_CHECK(cudaMemsetAsync(counts, 0, (size_t)(n_pred + 1) * (size_t)(n_gt + 1) * sizeof(unsigned long long), stream));
    if (n == 0) return TL_OK;
    TL_REQUIRE(pred && gt, "tl_cooccurrence_counts: null label pointer");
    k_cooccurrence<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(pred, gt, n, n_pred, n_gt, counts);
    TL_LAUNCH_CHECK();
    return TL_OK;
}

}  // extern "C"
