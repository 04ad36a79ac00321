// Geometry kernels: point->voxel (Morton-sorted), strided level maps, submanifold rulebook.
// All integer work here is bit-exact against oracle/spconv_ref.py (sets / partitions; row order is
// this library's own: Morton order, which the reference leaves implementation-defined).
#include <cub/cub.cuh>
#include <stdarg.h>

#include "tl_common.cuh"

namespace tl {

static thread_local char g_err[512] = "";
static long long g_launches = 0;
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches += n; }

// ------------------------------------------------------------------------------------------------
// point -> voxel
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int float_to_ordered(float f) {
    int v = __float_as_int(f);
    return v >= 0 ? v : v ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int v) { return __int_as_float(v >= 0 ? v : v ^ 0x7fffffff); }

__global__ void k_fill_i32(int* p, int v, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// per-batch-element minimum of the coordinates (tree_learn.py:134).  batch_ids is ascending, so a
// block whose first and last point share a batch id reduces in registers and issues 3 atomics.
__global__ void __launch_bounds__(256) k_batch_min(const float* __restrict__ coords, const int64_t* __restrict__ bids,
                                                   int64_t n, int* __restrict__ bmin) {
    const int64_t base = (int64_t)blockIdx.x * 1024;
    const int64_t end = min(base + 1024, n);
    const int b_first = (int)bids[base], b_last = (int)bids[end - 1];
    if (b_first == b_last) {
        float m[3] = {INFINITY, INFINITY, INFINITY};
        for (int64_t i = base + threadIdx.x; i < end; i += 256)
            for (int a = 0; a < 3; ++a) m[a] = fminf(m[a], coords[i * 3 + a]);
        __shared__ float red[3][8];
        for (int a = 0; a < 3; ++a) {
            float v = m[a];
            for (int o = 16; o; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
            if ((threadIdx.x & 31) == 0) red[a][threadIdx.x >> 5] = v;
        }
        __syncthreads();
        if (threadIdx.x < 3) {
            float v = red[threadIdx.x][0];
            for (int w = 1; w < 8; ++w) v = fminf(v, red[threadIdx.x][w]);
            atomicMin(&bmin[b_first * 3 + threadIdx.x], float_to_ordered(v));
        }
    } else {
        for (int64_t i = base + threadIdx.x; i < end; i += 256) {
            int b = (int)bids[i];
            for (int a = 0; a < 3; ++a) atomicMin(&bmin[b * 3 + a], float_to_ordered(coords[i * 3 + a]));
        }
    }
}

// voxel index c = floor((p - min) / vsize) evaluated in fp32 with IEEE division, like spconv's
// point2voxel kernel (oracle/spconv_ref.py voxel_index_fp32).
__global__ void k_point_keys(const float* __restrict__ coords, const int64_t* __restrict__ bids, int64_t n,
                             const int* __restrict__ bmin, float vsize, uint64_t* __restrict__ keys,
                             int* __restrict__ idx) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int b = (int)bids[i];
    uint32_t c[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float mn = ordered_to_float(bmin[b * 3 + a]);
        float v = floorf(__fdiv_rn(__fsub_rn(coords[i * 3 + a], mn), vsize));
        c[a] = (uint32_t)min(max((int)v, 0), (1 << kCoordBits) - 1);
    }
    keys[i] = make_key((uint32_t)b, c[0], c[1], c[2]);
    idx[i] = (int)i;
}

__global__ void k_mark_heads(const uint64_t* __restrict__ keys, int64_t n, int* __restrict__ flag) {
    int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= n) return;
    flag[j] = (j == 0 || keys[j] != keys[j - 1]) ? 1 : 0;
}

// one thread per sorted point: v2p for everyone, voxel record + mean of the first <=P points at heads
__global__ void k_emit_voxels(const uint64_t* __restrict__ skeys, const int* __restrict__ sidx,
                              const int* __restrict__ scan, int64_t n, const float* __restrict__ coords,
                              const float* __restrict__ feats, int n_feat, int use_coords, int use_feats, int max_pts,
                              uint64_t* __restrict__ vkeys, int* __restrict__ vcoords, float* __restrict__ vfeats,
                              int64_t* __restrict__ v2p) {
    int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint64_t key = skeys[j];
    const int vid = scan[j] - 1;
    v2p[sidx[j]] = vid;
    if (j != 0 && skeys[j - 1] == key) return;
    vkeys[vid] = key;
    int b, x, y, z;
    split_key(key, b, x, y, z);
    reinterpret_cast<int4*>(vcoords)[vid] = make_int4(b, x, y, z);
    const int nf = 3 + n_feat;
    float sum[3 + 8];
    for (int c = 0; c < nf; ++c) sum[c] = 0.f;
    int cnt = 0;
    for (int t = 0; t < max_pts && j + t < n && skeys[j + t] == key; ++t) {
        const int64_t i = sidx[j + t];
        float row[3 + 8];
        bool all_zero = true;
        for (int c = 0; c < 3; ++c) row[c] = coords[i * 3 + c];
        for (int c = 0; c < n_feat; ++c) row[3 + c] = feats[i * n_feat + c];
        for (int c = 0; c < nf; ++c) all_zero = all_zero && (row[c] == 0.f);
        if (!all_zero) {  // tree_learn.py:149-150: all-zero rows count as padding
            for (int c = 0; c < nf; ++c) sum[c] = __fadd_rn(sum[c], row[c]);
            ++cnt;
        }
    }
    float* out = vfeats + (int64_t)vid * nf;
    for (int c = 0; c < n_feat; ++c) out[c] = use_feats ? __fdiv_rn(sum[3 + c], (float)cnt) : 1.f;
    for (int c = 0; c < 3; ++c) out[n_feat + c] = use_coords ? __fdiv_rn(sum[c], (float)cnt) : 1.f;
}

struct VoxWs {
    int* bmin;
    uint64_t *keys_in, *keys_out;
    int *idx_in, *idx_out, *flag, *scan;
    void* cub_tmp;
    size_t cub_bytes;
};

static size_t cub_bytes_for(int64_t n) {
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (uint64_t*)nullptr, (uint64_t*)nullptr, (int*)nullptr, (int*)nullptr,
                                    (int)n);
    cub::DeviceScan::InclusiveSum(nullptr, b, (int*)nullptr, (int*)nullptr, (int)n);
    return align_up(a > b ? a : b);
}

static VoxWs carve_vox(void* ws, size_t bytes, int64_t n, int batch, bool& ok) {
    Carver c(ws, bytes);
    VoxWs w;
    w.bmin = c.take<int>((size_t)batch * 3 + 1);
    w.keys_in = c.take<uint64_t>(n);
    w.keys_out = c.take<uint64_t>(n);
    w.idx_in = c.take<int>(n);
    w.idx_out = c.take<int>(n);
    w.flag = c.take<int>(n);
    w.scan = c.take<int>(n);
    w.cub_bytes = cub_bytes_for(n);
    w.cub_tmp = c.take<char>(w.cub_bytes);
    ok = c.ok();
    return w;
}

// ------------------------------------------------------------------------------------------------
// strided (k=2, s=2) level map
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool parent_of(uint64_t key, int3 cshape, uint64_t& pkey) {
    int b, x, y, z;
    split_key(key, b, x, y, z);
    pkey = (key & ~kMortonMask) | ((key & kMortonMask) >> 3);
    return (x >> 1) < cshape.x && (y >> 1) < cshape.y && (z >> 1) < cshape.z;
}

__global__ void k_level_flags(const uint64_t* __restrict__ fkeys, int64_t n, int3 cshape, int* __restrict__ flag) {
    int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint64_t pk, pprev = kEmptyKey;
    bool valid = parent_of(fkeys[j], cshape, pk);
    if (j > 0) parent_of(fkeys[j - 1], cshape, pprev);
    // children of one parent share its coordinates => they are all kept or all dropped
    flag[j] = (valid && (j == 0 || pk != pprev)) ? 1 : 0;
}

__global__ void k_level_emit(const uint64_t* __restrict__ fkeys, const int* __restrict__ scan, int64_t n, int3 cshape,
                             int64_t down_stride, int64_t up_stride, uint64_t* __restrict__ ckeys,
                             int* __restrict__ ccoords, int* __restrict__ down_index, uint32_t* __restrict__ down_mask,
                             int* __restrict__ up_index, uint32_t* __restrict__ up_mask) {
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    // no early exit: the per-tile offset masks are OR-ed warp-wide first (one atomic per warp and mask word instead of one per
    // voxel: 128 voxels hit the same word)
    const uint64_t key = j < n ? fkeys[j] : 0;
    uint64_t pk = 0;
    const bool live = j < n && parent_of(key, cshape, pk);  // else dropped by the odd-edge rule (App. A.3): no pair
    const int par = live ? scan[j] - 1 : 0;
    const int kappa = (int)(key & 7);
    const unsigned bit = live ? 1u << kappa : 0u;
    const int lane = threadIdx.x & 31;
    const int dword = live ? par / TL_TILE_ROWS : -1;
    const unsigned grp = __match_any_sync(0xffffffffu, dword);
    const unsigned dbits = __reduce_or_sync(grp, bit);
    if (live && lane == __ffs(grp) - 1) atomicOr(&down_mask[dword], dbits);
    const unsigned ubits = __reduce_or_sync(0xffffffffu, bit);     // 32 consecutive rows share one 128-row tile
    if (lane == 0 && ubits) atomicOr(&up_mask[j / TL_TILE_ROWS], ubits);
    if (!live) return;
    down_index[kappa * down_stride + par] = (int)j;
    up_index[kappa * up_stride + j] = par;
    if (j == 0 || scan[j - 1] != scan[j]) {  // first child of this parent
        ckeys[par] = pk;
        int b, x, y, z;
        split_key(pk, b, x, y, z);
        reinterpret_cast<int4*>(ccoords)[par] = make_int4(b, x, y, z);
    }
}

// ------------------------------------------------------------------------------------------------
// submanifold rulebook
// ------------------------------------------------------------------------------------------------
// The rulebook's table keeps key and value in ONE 16 B slot {key, value, pad}: a probe that hits costs one 32 B L2 sector
// instead of two (the probe kernel is bound by random L2 sectors: 26 probes per voxel, 45 % hits; ncu: 469 us for the
// 1.77 M-voxel level at 36 % L2 / 7.6 % DRAM throughput -- issuing the probes of 9 offsets together made it slower, 539 us).
struct __align__(16) RbSlot {
    unsigned long long key;
    int val, pad;
};
__global__ void k_hash_build(const uint64_t* __restrict__ keys, int64_t n, RbSlot* table, uint64_t mask) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long key = keys[i];
    uint64_t slot = hash64(key) & mask;
    while (true) {
        const unsigned long long prev = atomicCAS(&table[slot].key, (unsigned long long)kEmptyKey, key);
        if (prev == kEmptyKey || prev == key) {
            table[slot].val = (int)i;
            return;
        }
        slot = (slot + 1) & mask;
    }
}
__device__ __forceinline__ int rb_find(const RbSlot* __restrict__ table, uint64_t mask, uint64_t key) {
    uint64_t slot = hash64(key) & mask;
    while (true) {
        const ulonglong2 s = __ldg(reinterpret_cast<const ulonglong2*>(table + slot));
        if (s.x == key) return (int)(uint32_t)s.y;
        if (s.x == kEmptyKey) return -1;
        slot = (slot + 1) & mask;
    }
}

__global__ void __launch_bounds__(TL_TILE_ROWS) k_subm_probe(const uint64_t* __restrict__ keys, int64_t n,
                                                             int64_t stride, int3 shape,
                                                             const RbSlot* __restrict__ table, uint64_t mask,
                                                             int* __restrict__ nbr, uint32_t* __restrict__ tile_mask) {
    __shared__ unsigned s_mask;
    if (threadIdx.x == 0) s_mask = 0;
    __syncthreads();
    const int64_t v = (int64_t)blockIdx.x * TL_TILE_ROWS + threadIdx.x;
    const bool live = v < n;
    uint64_t bbits = 0, px[3], py[3], pz[3];
    bool okx[3], oky[3], okz[3];
    if (live) {
        const uint64_t key = keys[v];
        int b, x, y, z;
        split_key(key, b, x, y, z);
        bbits = key & ~kMortonMask;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            int nx = x + d - 1, ny = y + d - 1, nz = z + d - 1;
            okx[d] = nx >= 0 && nx < shape.x;
            oky[d] = ny >= 0 && ny < shape.y;
            okz[d] = nz >= 0 && nz < shape.z;
            px[d] = part1by2((uint32_t)nx) << 2;
            py[d] = part1by2((uint32_t)ny) << 1;
            pz[d] = part1by2((uint32_t)nz);
        }
    }
    unsigned wmask = 0;
#pragma unroll
    for (int k = 0; k < 27; ++k) {
        const int a = k / 9, b2 = (k / 3) % 3, c = k % 3;
        int found = -1;
        if (live) {
            if (k == 13) found = (int)v;
            else if (okx[a] && oky[b2] && okz[c]) found = rb_find(table, mask, bbits | px[a] | py[b2] | pz[c]);
        }
        nbr[k * stride + v] = found;
        if (__any_sync(0xffffffffu, found >= 0)) wmask |= 1u << k;
    }
    if ((threadIdx.x & 31) == 0) atomicOr(&s_mask, wmask);
    __syncthreads();
    if (threadIdx.x == 0) tile_mask[blockIdx.x] = s_mask;
}

}  // namespace tl

using namespace tl;

extern "C" {

const char* tl_last_error(void) { return g_err; }
int tl_version(void) { return 1; }
long long tl_launch_count(void) { return g_launches; }
void tl_reset_launch_count(void) { g_launches = 0; }

size_t tl_voxelize_workspace_bytes(int64_t n) {
    if (n <= 0) return 256;
    return align_up(3 * 65536 * sizeof(int)) + 2 * align_up(n * 8) + 4 * align_up(n * 4) + cub_bytes_for(n) + 4096;
}

int tl_voxelize(const float* coords, const float* feats, int32_t n_feat, const int64_t* batch_ids, int64_t n,
                int32_t batch_size, float voxel_size, int32_t use_coords, int32_t use_feats, int32_t max_pts,
                uint64_t* voxel_keys, int32_t* voxel_coords, float* voxel_feats, int64_t* v2p, int64_t* num_voxels,
                void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    TL_REQUIRE(n > 0 && n < (1ll << 31), "tl_voxelize: n_points=%lld out of range", (long long)n);
    TL_REQUIRE(batch_size > 0 && batch_size < 32768, "tl_voxelize: batch_size=%d out of range", batch_size);
    TL_REQUIRE(n_feat >= 0 && n_feat <= 8, "tl_voxelize: n_feat=%d (max 8)", n_feat);
    bool ok;
    VoxWs w = carve_vox(workspace, workspace_bytes, n, batch_size, ok);
    TL_REQUIRE(ok, "tl_voxelize: workspace too small (%zu bytes)", workspace_bytes);
    const int T = 256;
    k_fill_i32<<<(batch_size * 3 + T - 1) / T, T, 0, stream>>>(w.bmin, 0x7fffffff, batch_size * 3);
    TL_LAUNCH_CHECK();
    k_batch_min<<<(unsigned)((n + 1023) / 1024), 256, 0, stream>>>(coords, batch_ids, n, w.bmin);
    TL_LAUNCH_CHECK();
    k_point_keys<<<(unsigned)((n + T - 1) / T), T, 0, stream>>>(coords, batch_ids, n, w.bmin, voxel_size, w.keys_in,
                                                                 w.idx_in);
    TL_LAUNCH_CHECK();
    int batch_bits = 0;
    while ((1 << batch_bits) < batch_size) ++batch_bits;
    size_t cb = w.cub_bytes;
    TL_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(w.cub_tmp, cb, w.keys_in, w.keys_out, w.idx_in, w.idx_out, (int)n, 0,
                                                  48 + batch_bits, stream));
    k_mark_heads<<<(unsigned)((n + T - 1) / T), T, 0, stream>>>(w.keys_out, n, w.flag);
    TL_LAUNCH_CHECK();
    cb = w.cub_bytes;
    TL_CUDA_CHECK(cub::DeviceScan::InclusiveSum(w.cub_tmp, cb, w.flag, w.scan, (int)n, stream));
    k_emit_voxels<<<(unsigned)((n + T - 1) / T), T, 0, stream>>>(w.keys_out, w.idx_out, w.scan, n, coords, feats, n_feat,
                                                                  use_coords, use_feats, max_pts, voxel_keys,
                                                                  voxel_coords, voxel_feats, v2p);
    TL_LAUNCH_CHECK();
    int m = 0;
    TL_CUDA_CHECK(cudaMemcpyAsync(&m, w.scan + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, stream));
    TL_CUDA_CHECK(cudaStreamSynchronize(stream));
    *num_voxels = m;
    return TL_OK;
}

size_t tl_level_workspace_bytes(int64_t n) {
    if (n <= 0) return 256;
    return 2 * align_up(n * 4) + cub_bytes_for(n) + 1024;
}

int tl_build_level(const uint64_t* fine_keys, int64_t n, const int32_t* fine_shape, uint64_t* coarse_keys,
                   int32_t* coarse_coords, int32_t* down_index, uint32_t* down_mask, int32_t* up_index,
                   uint32_t* up_mask, int32_t* coarse_shape, int64_t* n_coarse, void* workspace, size_t workspace_bytes,
                   void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    TL_REQUIRE(n > 0 && n < (1ll << 31), "tl_build_level: n_fine=%lld out of range", (long long)n);
    int3 cs;
    int* csp = &cs.x;
    for (int a = 0; a < 3; ++a) {
        csp[a] = (fine_shape[a] - 2) / 2 + 1;
        coarse_shape[a] = csp[a];
    }
    if (cs.x <= 0 || cs.y <= 0 || cs.z <= 0 || fine_shape[0] < 2 || fine_shape[1] < 2 || fine_shape[2] < 2) {
        set_error("your out spatial shape [%d, %d, %d] reach zero!!! input shape: [%d, %d, %d]", cs.x, cs.y, cs.z,
                  fine_shape[0], fine_shape[1], fine_shape[2]);
        return TL_ERR_REACH_ZERO;
    }
    Carver c(workspace, workspace_bytes);
    int* flag = c.take<int>(n);
    int* scan = c.take<int>(n);
    size_t cub_b = cub_bytes_for(n);
    void* cub_tmp = c.take<char>(cub_b);
    TL_REQUIRE(c.ok(), "tl_build_level: workspace too small");
    const int64_t stride = pad_rows(n);
    TL_CUDA_CHECK(cudaMemsetAsync(down_index, 0xFF, sizeof(int) * 8 * stride, stream));
    TL_CUDA_CHECK(cudaMemsetAsync(up_index, 0xFF, sizeof(int) * 8 * stride, stream));
    TL_CUDA_CHECK(cudaMemsetAsync(down_mask, 0, sizeof(uint32_t) * (stride / TL_TILE_ROWS), stream));
    TL_CUDA_CHECK(cudaMemsetAsync(up_mask, 0, sizeof(uint32_t) * (stride / TL_TILE_ROWS), stream));
    const int T = 256;
    k_level_flags<<<(unsigned)((n + T - 1) / T), T, 0, stream>>>(fine_keys, n, cs, flag);
    TL_LAUNCH_CHECK();
    TL_CUDA_CHECK(cub::DeviceScan::InclusiveSum(cub_tmp, cub_b, flag, scan, (int)n, stream));
    k_level_emit<<<(unsigned)((n + T - 1) / T), T, 0, stream>>>(fine_keys, scan, n, cs, stride, stride, coarse_keys,
                                                                 coarse_coords, down_index, down_mask, up_index, up_mask);
    TL_LAUNCH_CHECK();
    int m = 0;
    TL_CUDA_CHECK(cudaMemcpyAsync(&m, scan + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, stream));
    TL_CUDA_CHECK(cudaStreamSynchronize(stream));
    *n_coarse = m;
    return TL_OK;
}

size_t tl_rulebook_workspace_bytes(int64_t n) {
    uint64_t cap = table_capacity(n);
    return align_up(cap * 16) + 256;
}

int tl_subm_rulebook(const uint64_t* keys, int64_t n, const int32_t* spatial_shape, int32_t* nbr, uint32_t* tile_mask,
                     void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    TL_REQUIRE(n > 0 && n < (1ll << 31), "tl_subm_rulebook: n=%lld out of range", (long long)n);
    const uint64_t cap = table_capacity(n);
    Carver c(workspace, workspace_bytes);
    RbSlot* table = c.take<RbSlot>(cap);
    TL_REQUIRE(c.ok(), "tl_subm_rulebook: workspace too small");
    TL_CUDA_CHECK(cudaMemsetAsync(table, 0xFF, cap * sizeof(RbSlot), stream));      // key = kEmptyKey (all ones)
    const int T = 256;
    k_hash_build<<<(unsigned)((n + T - 1) / T), T, 0, stream>>>(keys, n, table, cap - 1);
    TL_LAUNCH_CHECK();
    const int64_t stride = pad_rows(n);
    int3 shape = make_int3(spatial_shape[0], spatial_shape[1], spatial_shape[2]);
    k_subm_probe<<<(unsigned)(stride / TL_TILE_ROWS), TL_TILE_ROWS, 0, stream>>>(keys, n, stride, shape, table, cap - 1, nbr,
                                                                                  tile_mask);
    TL_LAUNCH_CHECK();
    return TL_OK;
}

}  // extern "C"
