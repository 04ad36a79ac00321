// The step before the per-tile path: voxel down-sampling with trace (open3d
// VoxelDownSampleAndTrace) and the verticality feature (jakteristics radius-PCA).  Both bin the points into a uniform
// grid by one 63-bit radix sort; the down-sampler then reduces every cell in input order, the feature kernel scans the
// 27 cells around every point.  fp64 throughout, no FMA contraction where the result decides a cell or a neighbour.
#include <cub/cub.cuh>
#include <limits.h>

#include "tl_common.cuh"

namespace {

constexpr int kAxisBits = 21;                       // cells per axis after subtracting the per-axis minimum
constexpr long long kAxisCells = 1ll << kAxisBits;

__device__ __forceinline__ double round2(double v) { return __ddiv_rn(rint(__dmul_rn(v, 100.0)), 100.0); }

__device__ __forceinline__ long long cell_of(double v, double origin, double cell) {
    return (long long)floor(__ddiv_rn(__dsub_rn(v, origin), cell));
}

// per-axis minimum cell index (atomicMin over block minima)
__global__ void k_cell_min(const double* __restrict__ pts, int64_t n, int rnd, double origin, double cell,
                           long long* __restrict__ cmin) {
    long long m[3] = {LLONG_MAX, LLONG_MAX, LLONG_MAX};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const double v = rnd ? round2(pts[3 * i + d]) : pts[3 * i + d];
            const long long c = cell_of(v, origin, cell);
            m[d] = c < m[d] ? c : m[d];
        }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        for (int o = 16; o; o >>= 1) {
            const long long other = __shfl_xor_sync(0xffffffffu, m[d], o);
            m[d] = other < m[d] ? other : m[d];
        }
        if ((threadIdx.x & 31) == 0 && m[d] != LLONG_MAX) atomicMin(&cmin[d], m[d]);
    }
}

__global__ void k_cell_keys(const double* __restrict__ pts, int64_t n, int rnd, double origin, double cell,
                            const long long* __restrict__ cmin, uint64_t* __restrict__ keys, int* __restrict__ idx,
                            int* __restrict__ overflow) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t key = 0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const double v = rnd ? round2(pts[3 * i + d]) : pts[3 * i + d];
        const long long c = cell_of(v, origin, cell) - cmin[d];
        if (c < 0 || c >= kAxisCells) *overflow = 1;      // also catches NaN / inf coordinates
        key = (key << kAxisBits) | (uint64_t)(c & (kAxisCells - 1));
    }
    keys[i] = key;
    idx[i] = (int)i;
}

__global__ void k_heads(const uint64_t* __restrict__ keys, int64_t n, int* __restrict__ flag) {
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j < n) flag[j] = (j == 0 || keys[j] != keys[j - 1]) ? 1 : 0;
}

// One thread per sorted row: everyone writes its trace entry; the first row of a cell also walks the cell in input
// order (the radix sort is stable) and emits sum / count exactly as a sequential fp64 accumulation would.
__global__ void k_emit_cells(const uint64_t* __restrict__ skeys, const int* __restrict__ sidx, const int* __restrict__ scan,
                             int64_t n, const double* __restrict__ pts, int rnd, double* __restrict__ out_pts,
                             int64_t* __restrict__ first, int64_t* __restrict__ offsets, int64_t* __restrict__ trace) {
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= n) return;
    trace[j] = sidx[j];
    if (j == n - 1) offsets[scan[j]] = n;
    if (j > 0 && skeys[j] == skeys[j - 1]) return;
    const int v = scan[j] - 1;
    double s[3] = {0.0, 0.0, 0.0};
    int64_t e = j;
    for (; e < n && skeys[e] == skeys[j]; ++e) {
        const double* p = pts + 3 * (int64_t)sidx[e];
#pragma unroll
        for (int d = 0; d < 3; ++d) s[d] = __dadd_rn(s[d], rnd ? round2(p[d]) : p[d]);
    }
    const double cnt = (double)(e - j);
#pragma unroll
    for (int d = 0; d < 3; ++d) out_pts[3 * (int64_t)v + d] = __ddiv_rn(s[d], cnt);
    first[v] = sidx[j];
    offsets[v] = j;
}

__global__ void k_gather_sorted(const int* __restrict__ sidx, int64_t n, const double* __restrict__ pts,
                                double* __restrict__ sx, double* __restrict__ sy, double* __restrict__ sz) {
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= n) return;
    const double* p = pts + 3 * (int64_t)sidx[j];
    sx[j] = p[0], sy[j] = p[1], sz[j] = p[2];
}

__device__ __forceinline__ int64_t lower_bound(const uint64_t* __restrict__ a, int64_t n, uint64_t key) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Eigenvector of the smallest eigenvalue of a symmetric 3x3 matrix by cyclic Jacobi rotations (fp64); returns |v_z|.
__device__ double normal_abs_z(double a00, double a01, double a02, double a11, double a12, double a22) {
    double a[3][3] = {{a00, a01, a02}, {a01, a11, a12}, {a02, a12, a22}};
    double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 12; ++sweep) {
        const double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
        const double diag = fabs(a[0][0]) + fabs(a[1][1]) + fabs(a[2][2]);
        if (off <= 1e-300 || off <= 1e-17 * diag) break;
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int q = p + 1; q < 3; ++q) {
                if (a[p][q] == 0.0) continue;
                const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll
                for (int k = 0; k < 3; ++k) {      // A <- A J
                    const double akp = a[k][p], akq = a[k][q];
                    a[k][p] = c * akp - s * akq;
                    a[k][q] = s * akp + c * akq;
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) {      // A <- J^T A
                    const double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = c * apk - s * aqk;
                    a[q][k] = s * apk + c * aqk;
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) {      // V <- V J
                    const double vkp = v[k][p], vkq = v[k][q];
                    v[k][p] = c * vkp - s * vkq;
                    v[k][q] = s * vkp + c * vkq;
                }
            }
    }
    double lam = a[0][0], nx = v[0][0], ny = v[1][0], nz = v[2][0];
    if (a[1][1] < lam) lam = a[1][1], nx = v[0][1], ny = v[1][1], nz = v[2][1];
    if (a[2][2] < lam) lam = a[2][2], nx = v[0][2], ny = v[1][2], nz = v[2][2];
    return fabs(nz) / sqrt(nx * nx + ny * ny + nz * nz);
}

// One thread per cell-sorted point.  Neighbour cells that differ only in z are adjacent in key order, so the 27 cells
// are 9 contiguous runs of the sorted arrays; the lanes of a warp mostly share the runs (same or adjacent cells).
__global__ void __launch_bounds__(128) k_verticality(const uint64_t* __restrict__ skeys, const int* __restrict__ sidx,
                                                     const double* __restrict__ sx, const double* __restrict__ sy,
                                                     const double* __restrict__ sz, int64_t n, double r2,
                                                     double* __restrict__ out) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint64_t key = skeys[t];
    const long long cz = (long long)(key & (kAxisCells - 1)), cy = (long long)((key >> kAxisBits) & (kAxisCells - 1)),
                    cx = (long long)(key >> (2 * kAxisBits));
    const double qx = sx[t], qy = sy[t], qz = sz[t];
    double s1x = 0, s1y = 0, s1z = 0, sxx = 0, sxy = 0, sxz = 0, syy = 0, syz = 0, szz = 0;
    long long cnt = 0;
    for (long long ax = cx - 1; ax <= cx + 1; ++ax) {
        if (ax < 0 || ax >= kAxisCells) continue;
        for (long long ay = cy - 1; ay <= cy + 1; ++ay) {
            if (ay < 0 || ay >= kAxisCells) continue;
            const uint64_t base = ((uint64_t)ax << (2 * kAxisBits)) | ((uint64_t)ay << kAxisBits);
            const long long z0 = cz > 0 ? cz - 1 : 0, z1 = cz + 1 < kAxisCells ? cz + 1 : kAxisCells - 1;
            const int64_t lo = lower_bound(skeys, n, base | (uint64_t)z0);
            const int64_t hi = lower_bound(skeys, n, (base | (uint64_t)z1) + 1);
            for (int64_t j = lo; j < hi; ++j) {
                const double dx = __dsub_rn(sx[j], qx), dy = __dsub_rn(sy[j], qy), dz = __dsub_rn(sz[j], qz);
                const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                if (d2 <= r2) {
                    ++cnt;
                    s1x += dx, s1y += dy, s1z += dz;
                    sxx += dx * dx, sxy += dx * dy, sxz += dx * dz, syy += dy * dy, syz += dy * dz, szz += dz * dz;
                }
            }
        }
    }
    double res = __longlong_as_double(0x7ff8000000000000ll);        // NaN: fewer than 3 neighbours
    if (cnt >= 3) {
        const double inv = 1.0 / (double)cnt;
        res = 1.0 - normal_abs_z(sxx - s1x * s1x * inv, sxy - s1x * s1y * inv, sxz - s1x * s1z * inv,
                                 syy - s1y * s1y * inv, syz - s1y * s1z * inv, szz - s1z * s1z * inv);
    }
    out[sidx[t]] = res;
}

size_t cub_bytes_for(int64_t n) {
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (uint64_t*)nullptr, (uint64_t*)nullptr, (int*)nullptr, (int*)nullptr, (int)n);
    cub::DeviceScan::InclusiveSum(nullptr, b, (int*)nullptr, (int*)nullptr, (int)n);
    return tl::align_up(a > b ? a : b);
}

struct GridWs {
    long long* cmin;
    int* overflow;
    uint64_t *keys_in, *keys_out;
    int *idx_in, *idx_out, *flag, *scan;
    double *sx, *sy, *sz;
    void* cub_tmp;
    size_t cub_bytes;
    bool ok;
};

GridWs carve(void* ws, size_t bytes, int64_t n, bool sorted_points) {
    tl::Carver c(ws, bytes);
    GridWs w;
    w.cmin = c.take<long long>(4);
    w.overflow = c.take<int>(4);
    w.keys_in = c.take<uint64_t>(n);
    w.keys_out = c.take<uint64_t>(n);
    w.idx_in = c.take<int>(n);
    w.idx_out = c.take<int>(n);
    w.flag = c.take<int>(n);
    w.scan = c.take<int>(n);
    w.sx = w.sy = w.sz = nullptr;
    if (sorted_points) w.sx = c.take<double>(n), w.sy = c.take<double>(n), w.sz = c.take<double>(n);
    w.cub_bytes = cub_bytes_for(n);
    w.cub_tmp = c.take<char>(w.cub_bytes);
    w.ok = c.ok();
    return w;
}

size_t grid_ws_bytes(int64_t n, bool sorted_points) {
    if (n <= 0) return 256;
    const size_t a = tl::align_up((size_t)n * 8), b = tl::align_up((size_t)n * 4);
    return 2 * 256 + 2 * a + 4 * b + (sorted_points ? 3 * a : 0) + cub_bytes_for(n) + 1024;
}

// bins the points and sorts them by cell; returns TL_OK or an error (host sync for the overflow flag)
int sort_by_cell(const double* pts, int64_t n, int rnd, double origin, double cell, GridWs& w, cudaStream_t stream,
                 const char* who) {
    const int T = 256;
    const long long init[4] = {LLONG_MAX, LLONG_MAX, LLONG_MAX, 0};
    TL_CUDA_CHECK(cudaMemcpyAsync(w.cmin, init, sizeof(init), cudaMemcpyHostToDevice, stream));
    TL_CUDA_CHECK(cudaMemsetAsync(w.overflow, 0, sizeof(int), stream));
    const unsigned blocks = (unsigned)((n + T - 1) / T < 148 * 8 ? (n + T - 1) / T : 148 * 8);
    k_cell_min<<<blocks, T, 0, stream>>>(pts, n, rnd, origin, cell, w.cmin);
    TL_LAUNCH_CHECK();
    k_cell_keys<<<(unsigned)((n + T - 1) / T), T, 0, stream>>>(pts, n, rnd, origin, cell, w.cmin, w.keys_in, w.idx_in,
                                                               w.overflow);
    TL_LAUNCH_CHECK();
    size_t cb = w.cub_bytes;
    TL_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(w.cub_tmp, cb, w.keys_in, w.keys_out, w.idx_in, w.idx_out, (int)n, 0,
                                                  3 * kAxisBits, stream));
    tl::count_launch(3);
    int overflow = 0;
    TL_CUDA_CHECK(cudaMemcpyAsync(&overflow, w.overflow, sizeof(int), cudaMemcpyDeviceToHost, stream));
    TL_CUDA_CHECK(cudaStreamSynchronize(stream));
    TL_REQUIRE(!overflow, "%s: the points span more than 2^%d cells of %g along an axis (or hold NaN/inf)", who, kAxisBits, cell);
    return TL_OK;
}

}  // namespace

extern "C" {

size_t tl_downsample_workspace_bytes(int64_t n_points) { return grid_ws_bytes(n_points, false); }

int tl_voxel_downsample_trace(const double* points, int64_t n, int32_t round2_first, double voxel_size,
                              double voxel_min_bound, double* out_points, int64_t* first_index, int64_t* offsets,
                              int64_t* trace, int64_t* n_voxels, void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    TL_REQUIRE(n_voxels, "tl_voxel_downsample_trace: null n_voxels");
    *n_voxels = 0;
    TL_REQUIRE(n >= 0 && n < (1ll << 31), "tl_voxel_downsample_trace: n_points=%lld out of range", (long long)n);
    TL_REQUIRE(voxel_size > 0.0, "tl_voxel_downsample_trace: voxel_size <= 0.");       // open3d raises the same way
    TL_REQUIRE(offsets, "tl_voxel_downsample_trace: null offsets");
    if (n == 0) {
        TL_CUDA_CHECK(cudaMemsetAsync(offsets, 0, sizeof(int64_t), stream));
        return TL_OK;
    }
    TL_REQUIRE(points && out_points && first_index && trace, "tl_voxel_downsample_trace: null pointer");
    TL_REQUIRE(workspace && workspace_bytes >= grid_ws_bytes(n, false), "tl_voxel_downsample_trace: workspace too small");
    GridWs w = carve(workspace, workspace_bytes, n, false);
    TL_REQUIRE(w.ok, "tl_voxel_downsample_trace: workspace too small");
    int rc = sort_by_cell(points, n, round2_first, voxel_min_bound, voxel_size, w, stream, "tl_voxel_downsample_trace");
    if (rc != TL_OK) return rc;
    const int T = 256;
    k_heads<<<(unsigned)((n + T - 1) / T), T, 0, stream>>>(w.keys_out, n, w.flag);
    TL_LAUNCH_CHECK();
    size_t cb = w.cub_bytes;
    TL_CUDA_CHECK(cub::DeviceScan::InclusiveSum(w.cub_tmp, cb, w.flag, w.scan, (int)n, stream));
    tl::count_launch(2);
    k_emit_cells<<<(unsigned)((n + T - 1) / T), T, 0, stream>>>(w.keys_out, w.idx_out, w.scan, n, points, round2_first,
                                                                out_points, first_index, offsets, trace);
    TL_LAUNCH_CHECK();
    int m = 0;
    TL_CUDA_CHECK(cudaMemcpyAsync(&m, w.scan + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, stream));
    TL_CUDA_CHECK(cudaStreamSynchronize(stream));
    *n_voxels = m;
    return TL_OK;
}

size_t tl_verticality_workspace_bytes(int64_t n_points) { return grid_ws_bytes(n_points, true); }

int tl_verticality(const double* points, int64_t n, double search_radius, double* out, void* workspace,
                   size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    TL_REQUIRE(n >= 0 && n < (1ll << 31), "tl_verticality: n_points=%lld out of range", (long long)n);
    TL_REQUIRE(search_radius > 0.0, "tl_verticality: search_radius must be positive");
    if (n == 0) return TL_OK;
    TL_REQUIRE(points && out, "tl_verticality: null pointer");
    TL_REQUIRE(workspace && workspace_bytes >= grid_ws_bytes(n, true), "tl_verticality: workspace too small");
    GridWs w = carve(workspace, workspace_bytes, n, true);
    TL_REQUIRE(w.ok, "tl_verticality: workspace too small");
    // cells a shade wider than the radius: two points within the radius are then in adjacent cells even when the
    // rounding of coordinate / cell lands one of them on the other side of a cell boundary
    int rc = sort_by_cell(points, n, 0, 0.0, search_radius * 1.000001, w, stream, "tl_verticality");
    if (rc != TL_OK) return rc;
    const int T = 256;
    k_gather_sorted<<<(unsigned)((n + T - 1) / T), T, 0, stream>>>(w.idx_out, n, points, w.sx, w.sy, w.sz);
    TL_LAUNCH_CHECK();
    k_verticality<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(w.keys_out, w.idx_out, w.sx, w.sy, w.sz, n,
                                                                   search_radius * search_radius, out);
    TL_LAUNCH_CHECK();
    return TL_OK;
}

}  // extern "C"
