// Overlap merge, DBSCAN-equivalent clustering and kNN vote.
// Integer results (group structure, partition, label ids, votes) are bit-exact against
// oracle/cluster_ref.py, which is pinned to the reference's pandas / sklearn code.
// Distances are evaluated in fp64 on the float32 inputs without FMA contraction, the way
// sklearn's KD-tree evaluates its reduced distance (sum of squares in axis order).
#include <cub/cub.cuh>

#include <algorithm>
#include <vector>

#include "tl_common.cuh"

namespace tl {

// ------------------------------------------------------------------------------------------------
// overlap merge: group-by round(coords, 2), mean of the value columns, sorted by (x,y,z)
// ------------------------------------------------------------------------------------------------
constexpr int kAxisBits = 21;
constexpr int64_t kAxisBias = 1 << (kAxisBits - 1);

__global__ void k_merge_keys(const float* __restrict__ coords, int64_t n, uint64_t* __restrict__ keys,
                             int* __restrict__ idx) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t key = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        // pandas round(2) on a float32 column == rint(x * 100) / 100 evaluated in float32
        int64_t q = (int64_t)rintf(__fmul_rn(coords[i * 3 + a], 100.0f)) + kAxisBias;
        q = q < 0 ? 0 : (q >= (1 << kAxisBits) ? (1 << kAxisBits) - 1 : q);
        key = (key << kAxisBits) | (uint64_t)q;
    }
    keys[i] = key;
    idx[i] = (int)i;
}

__global__ void k_heads_u64(const uint64_t* __restrict__ keys, int64_t n, int* __restrict__ flag) {
    int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= n) return;
    flag[j] = (j == 0 || keys[j] != keys[j - 1]) ? 1 : 0;
}

// seg_start[g] = first sorted position of group g; seg_start[n_groups] = n
__global__ void k_segment_starts(const int* __restrict__ flag, const int* __restrict__ scan, int64_t n,
                                 int* __restrict__ seg_start) {
    int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= n) return;
    if (flag[j]) seg_start[scan[j] - 1] = (int)j;
    if (j == n - 1) seg_start[scan[j]] = (int)n;
}

__global__ void k_merge_emit(const uint64_t* __restrict__ skeys, const int* __restrict__ sidx,
                             const int* __restrict__ scan, const int* __restrict__ seg_start, int64_t n, int n_val,
                             const float* __restrict__ values, float* __restrict__ out_coords,
                             float* __restrict__ out_values, int* __restrict__ group_of_row, int n_groups) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < n) group_of_row[sidx[t]] = scan[t] - 1;
    const int64_t total = (int64_t)n_groups * (n_val + 3);
    for (int64_t e = t; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int g = (int)(e / (n_val + 3)), col = (int)(e % (n_val + 3));
        const int s = seg_start[g], end = seg_start[g + 1];
        if (col < 3) {
            const uint64_t key = skeys[s];
            const int64_t q = (int64_t)((key >> (kAxisBits * (2 - col))) & ((1u << kAxisBits) - 1)) - kAxisBias;
            out_coords[(int64_t)g * 3 + col] = __fdiv_rn((float)q, 100.0f);
        } else {
            double acc = 0.0;
            for (int j = s; j < end; ++j) acc += (double)values[(int64_t)sidx[j] * n_val + (col - 3)];
            out_values[(int64_t)g * n_val + (col - 3)] = (float)(acc / (double)(end - s));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// uniform-grid cell index shared by the clustering and kNN kernels
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t cell_key3(int64_t cx, int64_t cy, int64_t cz) {
    return ((uint64_t)(cx + kAxisBias) << (2 * kAxisBits)) | ((uint64_t)(cy + kAxisBias) << kAxisBits) |
           (uint64_t)(cz + kAxisBias);
}

template <int DIM>
__global__ void k_cell_keys(const float* __restrict__ pts, int64_t n, double inv_cell, uint64_t* __restrict__ keys,
                            int* __restrict__ idx) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t c[3] = {0, 0, 0};
#pragma unroll
    for (int a = 0; a < DIM; ++a) {
        int64_t v = (int64_t)floor((double)pts[i * DIM + a] * inv_cell);
        c[a] = v < -kAxisBias + 2 ? -kAxisBias + 2 : (v > kAxisBias - 3 ? kAxisBias - 3 : v);
    }
    keys[i] = cell_key3(c[0], c[1], c[2]);
    idx[i] = (int)i;
}

template <int DIM>
__global__ void k_cell_table(const uint64_t* __restrict__ skeys, const int* __restrict__ sidx,
                             const int* __restrict__ flag, const int* __restrict__ scan, int64_t n,
                             const float* __restrict__ pts, float* __restrict__ spts, uint64_t* tkeys, int* tvals,
                             uint64_t mask) {
    int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int i = sidx[j];
#pragma unroll
    for (int a = 0; a < DIM; ++a) spts[j * DIM + a] = pts[(int64_t)i * DIM + a];
    if (flag[j]) hash_insert(tkeys, tvals, mask, skeys[j], scan[j] - 1);
}

struct CellGrid {
    uint64_t *keys_in, *keys_out, *tkeys;
    int *idx_in, *idx_out, *flag, *scan, *seg_start, *tvals;
    float* spts;
    uint64_t cap;
    void* cub_tmp;
    size_t cub_bytes;
};

static size_t cub_sort_scan_bytes(int64_t n) {
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (uint64_t*)nullptr, (uint64_t*)nullptr, (int*)nullptr, (int*)nullptr,
                                    (int)n);
    cub::DeviceScan::InclusiveSum(nullptr, b, (int*)nullptr, (int*)nullptr, (int)n);
    return align_up(a > b ? a : b);
}

static size_t grid_bytes(int64_t n) {
    uint64_t cap = table_capacity(n);
    return 2 * align_up(n * 8) + 5 * align_up((n + 1) * 4) + align_up(n * 3 * 4) + align_up(cap * 8) +
           align_up(cap * 4) + cub_sort_scan_bytes(n) + 2048;
}

static bool carve_grid(Carver& c, int64_t n, CellGrid& g) {
    g.keys_in = c.take<uint64_t>(n);
    g.keys_out = c.take<uint64_t>(n);
    g.idx_in = c.take<int>(n + 1);
    g.idx_out = c.take<int>(n + 1);
    g.flag = c.take<int>(n + 1);
    g.scan = c.take<int>(n + 1);
    g.seg_start = c.take<int>(n + 1);
    g.spts = c.take<float>(n * 3);
    g.cap = table_capacity(n);
    g.tkeys = c.take<uint64_t>(g.cap);
    g.tvals = c.take<int>(g.cap);
    g.cub_bytes = cub_sort_scan_bytes(n);
    g.cub_tmp = c.take<char>(g.cub_bytes);
    return c.ok();
}

// sort points into cells, build (cell key -> segment id) table and the cell-sorted coordinate copy
template <int DIM>
static int build_grid(const float* pts, int64_t n, double cell, CellGrid& g, cudaStream_t stream) {
    const int T = 256;
    const unsigned nb = (unsigned)((n + T - 1) / T);
    k_cell_keys<DIM><<<nb, T, 0, stream>>>(pts, n, 1.0 / cell, g.keys_in, g.idx_in);
    TL_LAUNCH_CHECK();
    size_t cb = g.cub_bytes;
    TL_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(g.cub_tmp, cb, g.keys_in, g.keys_out, g.idx_in, g.idx_out, (int)n, 0,
                                                  3 * kAxisBits, stream));
    k_heads_u64<<<nb, T, 0, stream>>>(g.keys_out, n, g.flag);
    TL_LAUNCH_CHECK();
    cb = g.cub_bytes;
    TL_CUDA_CHECK(cub::DeviceScan::InclusiveSum(g.cub_tmp, cb, g.flag, g.scan, (int)n, stream));
    k_segment_starts<<<nb, T, 0, stream>>>(g.flag, g.scan, n, g.seg_start);
    TL_LAUNCH_CHECK();
    TL_CUDA_CHECK(cudaMemsetAsync(g.tkeys, 0xFF, g.cap * 8, stream));
    k_cell_table<DIM><<<nb, T, 0, stream>>>(g.keys_out, g.idx_out, g.flag, g.scan, n, pts, g.spts, g.tkeys, g.tvals,
                                             g.cap - 1);
    TL_LAUNCH_CHECK();
    return TL_OK;
}

// ------------------------------------------------------------------------------------------------
// DBSCAN(eps, min_samples=2) == connected components of the d<=eps graph (min-index union-find)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int uf_find(int* parent, int x) {
    int p = parent[x];
    while (p != x) {
        int gp = parent[p];
        if (gp != p) parent[x] = gp;  // path halving (benign race: only ever points closer to the root)
        x = p;
        p = gp;
    }
    return x;
}
__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
    while (true) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        const int hi = a > b ? a : b, lo = a > b ? b : a;
        if (atomicCAS(&parent[hi], hi, lo) == hi) return;  // hook the larger root under the smaller: root == min index
    }
}

__global__ void k_iota(int* p, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = (int)i;
}

// Grid DBSCAN for min_samples = 2.  Cells have side < eps / sqrt(2): all points of a cell are mutually within eps, so a
// cell is chained together without distance tests (k_cc_chain); two different cells need ONE witness pair within eps
// to merge, and cells already in the same component are skipped by comparing roots (k_cc_link).  Dense blobs (every
// point of a tree shifted onto its base) therefore cost O(points x 12 neighbour cells) instead of O(points^2).
__global__ void k_cc_chain(const uint64_t* __restrict__ skeys, const int* __restrict__ sidx, int64_t n, int* parent) {
    int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= n || j == 0) return;
    if (skeys[j] == skeys[j - 1]) uf_union(parent, sidx[j], sidx[j - 1]);
}

__global__ void k_cc_link(const uint64_t* __restrict__ skeys, const int* __restrict__ sidx,
                          const float* __restrict__ spts, const int* __restrict__ seg_start,
                          const uint64_t* __restrict__ tkeys, const int* __restrict__ tvals, uint64_t mask, int64_t n,
                          double r2, int* parent) {
    int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int i = sidx[j];
    const double x = spts[j * 2], y = spts[j * 2 + 1];
    const uint64_t key = skeys[j];
    const int64_t cx = (int64_t)(key >> (2 * kAxisBits)) - kAxisBias;
    const int64_t cy = (int64_t)((key >> kAxisBits) & ((1u << kAxisBits) - 1)) - kAxisBias;
    // eps / cell < 2 => partners live at most 2 cells away; each unordered cell pair is visited from its "lower" cell
    for (int dx = 0; dx <= 2; ++dx)
        for (int dy = (dx == 0 ? 1 : -2); dy <= 2; ++dy) {
            const int seg = hash_find(tkeys, tvals, mask, cell_key3(cx + dx, cy + dy, 0));
            if (seg < 0) continue;
            const int s = seg_start[seg], e = seg_start[seg + 1];
            if (uf_find(parent, sidx[s]) == uf_find(parent, i)) continue;   // cells already connected
            for (int j2 = s; j2 < e; ++j2) {
                const double ddx = x - (double)spts[j2 * 2], ddy = y - (double)spts[j2 * 2 + 1];
                const double d2 = __dadd_rn(__dmul_rn(ddx, ddx), __dmul_rn(ddy, ddy));
                if (d2 <= r2) {
                    uf_union(parent, i, sidx[j2]);
                    break;   // one witness pair connects the two cells
                }
            }
        }
}

__global__ void k_cc_roots(int* parent, int64_t n, int* __restrict__ root, int* __restrict__ size) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const bool live = i < n;
    const int r = live ? uf_find(parent, (int)i) : -1;
    if (live) root[i] = r;
    // component sizes: lanes with the same root add once (a few huge components otherwise serialise on one address)
    const unsigned alive = __ballot_sync(0xffffffffu, live);
    if (live) {
        const unsigned peers = __match_any_sync(alive, r);
        if ((threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&size[r], __popc(peers));
    }
}

__global__ void k_cc_valid(const int* __restrict__ root, const int* __restrict__ size, int64_t n, int min_size,
                           int* __restrict__ valid) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    valid[i] = (root[i] == (int)i && size[i] >= 2 && size[i] >= min_size) ? 1 : 0;
}

__global__ void k_cc_labels(const int* __restrict__ root, const int* __restrict__ valid, const int* __restrict__ rank,
                            int64_t n, int64_t not_assigned, int64_t start_num, int64_t* __restrict__ labels) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int r = root[i];
    labels[i] = valid[r] ? start_num + (int64_t)(rank[r] - 1) : not_assigned;
}

// ------------------------------------------------------------------------------------------------
// kNN majority vote over a 3-D cell grid, expanding cube search + exact brute-force fallback
// ------------------------------------------------------------------------------------------------
// kNN vote.  One cell-sort serves three grid resolutions: the 48-bit key of a reference point is hierarchical,
//   [ 12-bit 4 m cell x,y,z | 2-bit 1 m sub-cell x,y,z | 2-bit 0.25 m sub-cell x,y,z ],
// so points sorted by it are grouped by 0.25 m cell, by 1 m cell (key >> 6) and by 4 m cell (key >> 12) at once;
// each level only adds a (cell prefix -> [first, last) row) hash table.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxK = 8;
constexpr int kKnnLevels = 3;
constexpr double kKnnCell0 = 0.25;      // level l cells are 0.25 * 4^l m
constexpr int kFineBits = 16;           // fine (0.25 m) cell index bits per axis (+-8 km around the origin)
constexpr int kFineBias = 1 << (kFineBits - 1);

__device__ __forceinline__ int knn_fine_index(float v) {
    int64_t c = (int64_t)floor((double)v * (1.0 / kKnnCell0)) + kFineBias;
    return (int)(c < 8 ? 8 : (c > (1 << kFineBits) - 9 ? (1 << kFineBits) - 9 : c));   // clamp, leaving room for the rings
}
// prefix key of the level-l cell with (biased) level-l cell indices (cx,cy,cz)
__device__ __forceinline__ uint64_t knn_prefix(int cx, int cy, int cz, int level) {
    uint64_t key = 0;
    const int top = kFineBits - 2 * level;     // bits per axis at this level
    const int coarse = top - 2 * (2 - level);  // bits of the 4 m part
    key = ((uint64_t)(cx >> (top - coarse)) << (2 * coarse)) | ((uint64_t)(cy >> (top - coarse)) << coarse) |
          (uint64_t)(cz >> (top - coarse));
    for (int sub = 2 - level - 1; sub >= 0; --sub)   // 2-bit digits below the 4 m part, most significant first
        key = (key << 6) | ((uint64_t)((cx >> (2 * sub)) & 3) << 4) | ((uint64_t)((cy >> (2 * sub)) & 3) << 2) |
              (uint64_t)((cz >> (2 * sub)) & 3);
    return key;
}

__global__ void k_knn_keys(const float* __restrict__ pts, int64_t n, uint64_t* __restrict__ keys, int* __restrict__ idx) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = knn_prefix(knn_fine_index(pts[i * 3]), knn_fine_index(pts[i * 3 + 1]), knn_fine_index(pts[i * 3 + 2]), 0);
    idx[i] = (int)i;
}

// cell-sorted coordinate copy with the original reference index in .w
__global__ void k_knn_gather(const int* __restrict__ sidx, int64_t n, const float* __restrict__ pts,
                             float4* __restrict__ spts) {
    int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int i = sidx[j];
    spts[j] = make_float4(pts[(int64_t)i * 3], pts[(int64_t)i * 3 + 1], pts[(int64_t)i * 3 + 2], __int_as_float(i));
}

// level table: first row of every run of equal (key >> shift); the run ends where the next run starts
__global__ void k_knn_table(const uint64_t* __restrict__ skeys, int64_t n, int shift, uint64_t* tkeys, int* tvals,
                            uint64_t mask) {
    int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint64_t key = skeys[j] >> shift;
    if (j == 0 || (skeys[j - 1] >> shift) != key) hash_insert(tkeys, tvals, mask, key, (int)j);
}

struct TopK {
    double d[kMaxK];
    int idx[kMaxK];
    int k;
    float worst_f;   // fp32 upper bound of d[k-1] (pre-filter threshold), kept in a register
    __device__ void init(int k_) {
        k = k_;
        for (int t = 0; t < kMaxK; ++t) d[t] = 1e300, idx[t] = 0x7fffffff;
        worst_f = __int_as_float(0x7f800000);
    }
    // ordered by (distance, reference index): deterministic under ties
    __device__ void push(double dist, int index) {
        if (dist > d[k - 1] || (dist == d[k - 1] && index >= idx[k - 1])) return;
        int t = k - 1;
        while (t > 0 && (d[t - 1] > dist || (d[t - 1] == dist && idx[t - 1] > index))) {
            d[t] = d[t - 1], idx[t] = idx[t - 1];
            --t;
        }
        d[t] = dist, idx[t] = index;
        worst_f = (float)d[k - 1] * 1.00001f + 1e-30f;
    }
    __device__ int64_t vote(const int64_t* __restrict__ labels) const {  // most frequent label, ties -> smallest label
        int64_t lab[kMaxK];
        for (int a = 0; a < k; ++a) lab[a] = labels[idx[a]];
        int64_t best = 0;
        int best_cnt = 0;
        for (int a = 0; a < k; ++a) {
            int cnt = 0;
            for (int b = 0; b < k; ++b) cnt += (lab[b] == lab[a]);
            if (cnt > best_cnt || (cnt == best_cnt && lab[a] < best)) best = lab[a], best_cnt = cnt;
        }
        return best;
    }
};

struct KnnLevel {
    const uint64_t* tkeys;
    const int* tvals;
    uint64_t mask;
    int shift;      // key >> shift = this level's cell prefix
    int max_ring;
};

struct KnnGrid {
    const uint64_t* skeys;   // sorted hierarchical keys
    const float4* spts;      // cell-sorted (x, y, z, original index)
    int64_t n;
    KnnLevel level[kKnnLevels];
};

__device__ __forceinline__ void knn_scan_cell(const KnnGrid& g, const KnnLevel& lv, uint64_t prefix, float qx, float qy,
                                              float qz, TopK& top) {
    int j = hash_find(lv.tkeys, lv.tvals, lv.mask, prefix);
    if (j < 0) return;
    const double x = qx, y = qy, z = qz;
    for (; j < g.n; ++j) {
        if (lv.shift ? ((__ldg(g.skeys + j) >> lv.shift) != prefix) : (__ldg(g.skeys + j) != prefix)) break;
        const float4 p = __ldg(g.spts + j);
        // fp32 pre-filter (relative error of the fp32 sum of squares << 1e-5), then the exact fp64 distance
        const float fx = qx - p.x, fy = qy - p.y, fz = qz - p.z;
        const float d2f = fx * fx + fy * fy + fz * fz;
        if (d2f > top.worst_f) continue;
        const double ax = x - (double)p.x, ay = y - (double)p.y, az = z - (double)p.z;
        const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)), __dmul_rn(az, az));
        top.push(d2, __float_as_int(p.w));
    }
}

// Multi-resolution search: level 0 has small cells (dense tree-base blobs after the offset shift), coarser levels catch
// sparse neighbourhoods without probing hundreds of empty cells; exact scan of all references as the last resort.
// Within a level the rings are scanned shell by shell (nothing is re-read); a level's result is final once the k-th
// distance is within the radius that level has provably covered.
__global__ void __launch_bounds__(128) k_knn_vote(const float* __restrict__ query, int64_t nq, const KnnGrid g,
                                                  const int64_t* __restrict__ ref_labels, int k,
                                                  const int* __restrict__ order, int64_t* __restrict__ out) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nq) return;
    const int64_t q = order ? order[t] : t;     // queries in cell order: the lanes of a warp scan the same cells
    const float qx = query[q * 3], qy = query[q * 3 + 1], qz = query[q * 3 + 2];
    const int fx = knn_fine_index(qx), fy = knn_fine_index(qy), fz = knn_fine_index(qz);
    TopK top;
    bool done = false;
    for (int level = 0; level < kKnnLevels && !done; ++level) {
        const KnnLevel& lv = g.level[level];
        const int cx = fx >> (2 * level), cy = fy >> (2 * level), cz = fz >> (2 * level);
        const double cell = kKnnCell0 * (double)(1 << (2 * level));
        top.init(k);
        for (int ring = 0; ring <= lv.max_ring && !done; ++ring) {
            for (int dx = -ring; dx <= ring; ++dx)
                for (int dy = -ring; dy <= ring; ++dy) {
                    const bool edge = (dx == -ring || dx == ring || dy == -ring || dy == ring);
                    for (int dz = -ring; dz <= ring; dz += (edge || ring == 0) ? 1 : 2 * ring)   // shell cells only
                        knn_scan_cell(g, lv, knn_prefix(cx + dx, cy + dy, cz + dz, level), qx, qy, qz, top);
                }
            // every reference closer than ring*cell was seen (the query lies inside the centre cell)
            const double safe = (double)ring * cell;
            done = ring > 0 && top.d[k - 1] <= safe * safe;
        }
    }
    if (!done) {  // farther than every grid reaches: exact scan of all references (rare)
        top.init(k);
        const double x = qx, y = qy, z = qz;
        for (int64_t j = 0; j < g.n; ++j) {
            const float4 p = __ldg(g.spts + j);
            const double ax = x - (double)p.x, ay = y - (double)p.y, az = z - (double)p.z;
            const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)), __dmul_rn(az, az));
            top.push(d2, __float_as_int(p.w));
        }
    }
    out[q] = top.vote(ref_labels);
}

// ------------------------------------------------------------------------------------------------
// HDBSCAN (sklearn.cluster.HDBSCAN(min_cluster_size = m), the reference's default clusterer:
// tree_learn/util/pipeline.py:184-191).  sklearn 1.5-1.9 runs, for Euclidean input:
//   core distance = distance to the min_samples-th nearest neighbour, the point itself included (KD-tree query),
//   MST of the mutual-reachability graph max(core_a, core_b, d_ab) by Prim's algorithm over the data matrix
//   (O(n^2), started at node 0, strict '<' updates, first minimal index wins),
//   edges sorted by weight -> single-linkage dendrogram -> condensed tree -> stability -> excess-of-mass selection.
// Here: core distances by a cell-grid search, Prim as one persistent kernel (every step = a parallel relaxation +
// lexicographic (value, index) argmin + one grid barrier), both in fp64 without FMA contraction like the rest of this file;
// the dendrogram / condensed-tree pass (O(n), pointer chasing) runs on the host inside this library.
// ------------------------------------------------------------------------------------------------
constexpr int kCoreMaxK = 128;

static inline double __longlong_as_double_host(long long v) {
    double d;
    memcpy(&d, &v, sizeof(d));
    return d;
}
__global__ void k_fill_f64(double* p, double v, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void k_fill_i32b(int* p, int v, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__global__ void __launch_bounds__(128) k_core_dist(const uint64_t* __restrict__ skeys, const int* __restrict__ sidx,
                                                   const float* __restrict__ spts, const int* __restrict__ seg_start,
                                                   const uint64_t* __restrict__ tkeys, const int* __restrict__ tvals,
                                                   uint64_t mask, int64_t n, int k, double cell, int max_ring,
                                                   double* __restrict__ core) {
    int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= n) return;
    const double x = spts[j * 2], y = spts[j * 2 + 1];
    const uint64_t key = skeys[j];
    const int64_t cx = (int64_t)(key >> (2 * kAxisBits)) - kAxisBias;
    const int64_t cy = (int64_t)((key >> kAxisBits) & ((1u << kAxisBits) - 1)) - kAxisBias;
    double top[kCoreMaxK];   // k smallest squared distances, ascending (local memory)
    for (int t = 0; t < k; ++t) top[t] = 1e300;
    auto push = [&](double d2) {
        if (d2 >= top[k - 1]) return;
        int t = k - 1;
        while (t > 0 && top[t - 1] > d2) top[t] = top[t - 1], --t;
        top[t] = d2;
    };
    bool done = false;
    for (int ring = 0; ring <= max_ring && !done; ++ring) {
        for (int dx = -ring; dx <= ring; ++dx) {
            const bool edge = (dx == -ring || dx == ring);
            for (int dy = -ring; dy <= ring; dy += (edge || ring == 0) ? 1 : 2 * ring) {   // shell cells only
                const int seg = hash_find(tkeys, tvals, mask, cell_key3(cx + dx, cy + dy, 0));
                if (seg < 0) continue;
                for (int j2 = seg_start[seg]; j2 < seg_start[seg + 1]; ++j2) {
                    const double ax = x - (double)spts[j2 * 2], ay = y - (double)spts[j2 * 2 + 1];
                    push(__dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)));
                }
            }
        }
        const double safe = (double)ring * cell;   // every point closer than ring*cell has been seen
        done = ring > 0 && top[k - 1] <= safe * safe;
    }
    if (!done) {   // sparse neighbourhood: exact scan of all points
        for (int t = 0; t < k; ++t) top[t] = 1e300;
        for (int64_t j2 = 0; j2 < n; ++j2) {
            const double ax = x - (double)spts[j2 * 2], ay = y - (double)spts[j2 * 2 + 1];
            push(__dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)));
        }
    }
    core[sidx[j]] = __dsqrt_rn(top[k - 1]);
}

// all CTAs of the (co-resident, cooperatively launched) grid meet; `target` = number of arrivals expected so far
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        while (*(volatile unsigned*)counter < target) {
        }
        __threadfence();
    }
    __syncthreads();
}

struct PrimBest {
    double v;
    int j;
};
__device__ __forceinline__ PrimBest prim_min(PrimBest a, PrimBest b) {   // lexicographic (value, index)
    return (b.v < a.v || (b.v == a.v && b.j < a.j)) ? b : a;
}
__device__ __forceinline__ PrimBest prim_block_min(PrimBest b, PrimBest* sh) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        PrimBest o;
        o.v = __shfl_down_sync(0xffffffffu, b.v, off);
        o.j = __shfl_down_sync(0xffffffffu, b.j, off);
        b = prim_min(b, o);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    if (lane == 0) sh[warp] = b;
    __syncthreads();
    if (warp == 0) {
        b = lane < nw ? sh[lane] : PrimBest{1.7976931348623157e308, 0x7fffffff};
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            PrimBest o;
            o.v = __shfl_down_sync(0xffffffffu, b.v, off);
            o.j = __shfl_down_sync(0xffffffffu, b.j, off);
            b = prim_min(b, o);
        }
        if (lane == 0) sh[0] = b;
    }
    __syncthreads();
    b = sh[0];
    __syncthreads();
    return b;
}

__global__ void __launch_bounds__(512) k_prim(const float* __restrict__ pts, const double* __restrict__ core, int n,
                                              double* min_reach, int* source, unsigned char* in_tree, double* blk_v,
                                              int* blk_j, unsigned* counter, int* __restrict__ mst_src,
                                              int* __restrict__ mst_dst, double* __restrict__ mst_w) {
    __shared__ PrimBest sh[32];
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
    int cur = 0, par = 0;
    for (int step = 0; step < n - 1; ++step) {
        const double cx = pts[2 * cur], cy = pts[2 * cur + 1], cc = core[cur];
        PrimBest best{1.7976931348623157e308, 0x7fffffff};
        for (int j = gtid; j < n; j += gsize) {
            if (j == cur) {
                in_tree[j] = 1;
                continue;
            }
            if (in_tree[j]) continue;
            const double ax = cx - (double)pts[2 * j], ay = cy - (double)pts[2 * j + 1];
            const double d = __dsqrt_rn(__dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)));
            const double mrd = fmax(fmax(cc, core[j]), d);
            double mr = min_reach[j];
            if (mrd < mr) {
                mr = mrd;
                min_reach[j] = mrd;
                source[j] = cur;
            }
            best = prim_min(best, PrimBest{mr, j});
        }
        best = prim_block_min(best, sh);
        if (threadIdx.x == 0) blk_v[par * gridDim.x + blockIdx.x] = best.v, blk_j[par * gridDim.x + blockIdx.x] = best.j;
        grid_barrier(counter, (unsigned)(step + 1) * gridDim.x);
        PrimBest b{1.7976931348623157e308, 0x7fffffff};
        for (int t = threadIdx.x; t < (int)gridDim.x; t += blockDim.x)
            b = prim_min(b, PrimBest{__ldcg(&blk_v[par * gridDim.x + t]), __ldcg(&blk_j[par * gridDim.x + t])});
        b = prim_block_min(b, sh);
        if (gtid == 0) mst_src[step] = __ldcg(&source[b.j]), mst_dst[step] = b.j, mst_w[step] = b.v;
        cur = b.j;
        par ^= 1;
    }
}

}  // namespace tl

using namespace tl;

extern "C" {

size_t tl_merge_workspace_bytes(int64_t n) {
    if (n <= 0) return 256;
    return 2 * align_up(n * 8) + 5 * align_up((n + 1) * 4) + cub_sort_scan_bytes(n) + 2048;
}

int tl_merge_groupby_mean(const float* coords, const float* values, int64_t n, int32_t n_val, float* out_coords,
                          float* out_values, int32_t* group_of_row, int64_t* n_groups, void* workspace,
                          size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    TL_REQUIRE(n > 0 && n < (1ll << 31), "tl_merge_groupby_mean: n=%lld out of range", (long long)n);
    Carver c(workspace, workspace_bytes);
    uint64_t* keys_in = c.take<uint64_t>(n);
    uint64_t* keys_out = c.take<uint64_t>(n);
    int* idx_in = c.take<int>(n + 1);
    int* idx_out = c.take<int>(n + 1);
    int* flag = c.take<int>(n + 1);
    int* scan = c.take<int>(n + 1);
    int* seg_start = c.take<int>(n + 1);
    size_t cub_b = cub_sort_scan_bytes(n);
    void* cub_tmp = c.take<char>(cub_b);
    TL_REQUIRE(c.ok(), "tl_merge_groupby_mean: workspace too small");
    const int T = 256;
    const unsigned nb = (unsigned)((n + T - 1) / T);
    k_merge_keys<<<nb, T, 0, stream>>>(coords, n, keys_in, idx_in);
    TL_LAUNCH_CHECK();
    size_t cb = cub_b;
    TL_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(cub_tmp, cb, keys_in, keys_out, idx_in, idx_out, (int)n, 0,
                                                  3 * kAxisBits, stream));
    k_heads_u64<<<nb, T, 0, stream>>>(keys_out, n, flag);
    TL_LAUNCH_CHECK();
    cb = cub_b;
    TL_CUDA_CHECK(cub::DeviceScan::InclusiveSum(cub_tmp, cb, flag, scan, (int)n, stream));
    k_segment_starts<<<nb, T, 0, stream>>>(flag, scan, n, seg_start);
    TL_LAUNCH_CHECK();
    int ng = 0;
    TL_CUDA_CHECK(cudaMemcpyAsync(&ng, scan + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, stream));
    TL_CUDA_CHECK(cudaStreamSynchronize(stream));
    k_merge_emit<<<nb, T, 0, stream>>>(keys_out, idx_out, scan, seg_start, n, n_val, values, out_coords, out_values,
                                       group_of_row, ng);
    TL_LAUNCH_CHECK();
    *n_groups = ng;
    return TL_OK;
}

size_t tl_cluster_workspace_bytes(int64_t n) {
    if (n <= 0) return 256;
    return grid_bytes(n) + 5 * align_up((n + 1) * 4) + 1024;
}

int tl_cluster_radius_cc(const float* points_xy, int64_t n, double radius, int64_t min_cluster_size,
                         int64_t not_assigned_label, int64_t start_num, int64_t* labels, int64_t* n_clusters,
                         void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    *n_clusters = 0;
    if (n == 0) return TL_OK;
    TL_REQUIRE(n > 0 && n < (1ll << 31) && radius > 0, "tl_cluster_radius_cc: n=%lld radius=%g", (long long)n, radius);
    Carver c(workspace, workspace_bytes);
    CellGrid g;
    bool ok = carve_grid(c, n, g);
    int* parent = c.take<int>(n + 1);
    int* root = c.take<int>(n + 1);
    int* size = c.take<int>(n + 1);
    int* valid = c.take<int>(n + 1);
    int* rank = c.take<int>(n + 1);
    TL_REQUIRE(ok && c.ok(), "tl_cluster_radius_cc: workspace too small");
    // cell side a hair below eps / sqrt(2): the cell diagonal stays below eps whatever the rounding of floor(x / cell)
    int rc = build_grid<2>(points_xy, n, radius * 0.70710678118654752440 * (1.0 - 1e-7), g, stream);
    if (rc != TL_OK) return rc;
    const int T = 256;
    const unsigned nb = (unsigned)((n + T - 1) / T);
    k_iota<<<nb, T, 0, stream>>>(parent, n);
    TL_LAUNCH_CHECK();
    TL_CUDA_CHECK(cudaMemsetAsync(size, 0, sizeof(int) * n, stream));
    k_cc_chain<<<nb, T, 0, stream>>>(g.keys_out, g.idx_out, n, parent);
    TL_LAUNCH_CHECK();
    k_cc_link<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(g.keys_out, g.idx_out, g.spts, g.seg_start, g.tkeys,
                                                               g.tvals, g.cap - 1, n, radius * radius, parent);
    TL_LAUNCH_CHECK();
    k_cc_roots<<<nb, T, 0, stream>>>(parent, n, root, size);
    TL_LAUNCH_CHECK();
    const int min_size = (int)(min_cluster_size > 0x7fffffff ? 0x7fffffff : (min_cluster_size < 0 ? 0 : min_cluster_size));
    k_cc_valid<<<nb, T, 0, stream>>>(root, size, n, min_size, valid);
    TL_LAUNCH_CHECK();
    size_t cb = g.cub_bytes;
    TL_CUDA_CHECK(cub::DeviceScan::InclusiveSum(g.cub_tmp, cb, valid, rank, (int)n, stream));
    k_cc_labels<<<nb, T, 0, stream>>>(root, valid, rank, n, not_assigned_label, start_num, labels);
    TL_LAUNCH_CHECK();
    int nc = 0;
    TL_CUDA_CHECK(cudaMemcpyAsync(&nc, rank + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, stream));
    TL_CUDA_CHECK(cudaStreamSynchronize(stream));
    *n_clusters = nc;
    return TL_OK;
}

size_t tl_knn_workspace_bytes(int64_t n_ref, int64_t n_query) {
    if (n_ref <= 0) return 256;
    const uint64_t cap = table_capacity(n_ref);
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (uint64_t*)nullptr, (uint64_t*)nullptr, (int*)nullptr, (int*)nullptr,
                                    (int)n_ref);
    size_t cub_q = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_q, (uint64_t*)nullptr, (uint64_t*)nullptr, (int*)nullptr, (int*)nullptr,
                                    (int)(n_query > 0 ? n_query : 1));
    return 2 * align_up(n_ref * 8) + 2 * align_up(n_ref * 4) + align_up(n_ref * 16) +
           kKnnLevels * (align_up(cap * 8) + align_up(cap * 4)) + align_up(cub_bytes) + align_up(cub_q) +
           2 * align_up(n_query * 8) + 2 * align_up(n_query * 4) + 8192;
}

int tl_knn_vote(const float* ref_xyz, const int64_t* ref_labels, int64_t n_ref, const float* query_xyz, int64_t n_query,
                int32_t k, int64_t* out_labels, void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_query == 0) return TL_OK;
    TL_REQUIRE(k >= 1 && k <= kMaxK, "tl_knn_vote: k=%d (1..%d)", k, kMaxK);
    TL_REQUIRE(n_ref >= k && n_ref < (1ll << 31), "tl_knn_vote: n_ref=%lld must be >= k=%d", (long long)n_ref, k);
    Carver c(workspace, workspace_bytes);
    uint64_t* keys_in = c.take<uint64_t>(n_ref);
    uint64_t* keys_out = c.take<uint64_t>(n_ref);
    int* idx_in = c.take<int>(n_ref);
    int* idx_out = c.take<int>(n_ref);
    float4* spts = c.take<float4>(n_ref);
    const uint64_t cap = table_capacity(n_ref);
    uint64_t* tkeys[kKnnLevels];
    int* tvals[kKnnLevels];
    for (int l = 0; l < kKnnLevels; ++l) tkeys[l] = c.take<uint64_t>(cap), tvals[l] = c.take<int>(cap);
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (uint64_t*)nullptr, (uint64_t*)nullptr, (int*)nullptr, (int*)nullptr,
                                    (int)n_ref);
    void* cub_tmp = c.take<char>(cub_bytes);
    TL_REQUIRE(c.ok(), "tl_knn_vote: workspace too small");
    const int T = 256;
    const unsigned nb = (unsigned)((n_ref + T - 1) / T);
    k_knn_keys<<<nb, T, 0, stream>>>(ref_xyz, n_ref, keys_in, idx_in);
    TL_LAUNCH_CHECK();
    TL_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, keys_in, keys_out, idx_in, idx_out, (int)n_ref, 0,
                                                  3 * kFineBits, stream));
    k_knn_gather<<<nb, T, 0, stream>>>(idx_out, n_ref, ref_xyz, spts);
    TL_LAUNCH_CHECK();
    const int rings[kKnnLevels] = {2, 3, 4};
    KnnGrid g;
    g.skeys = keys_out, g.spts = spts, g.n = n_ref;
    for (int l = 0; l < kKnnLevels; ++l) {
        TL_CUDA_CHECK(cudaMemsetAsync(tkeys[l], 0xFF, cap * 8, stream));
        k_knn_table<<<nb, T, 0, stream>>>(keys_out, n_ref, 6 * l, tkeys[l], tvals[l], cap - 1);
        TL_LAUNCH_CHECK();
        g.level[l] = KnnLevel{tkeys[l], tvals[l], cap - 1, 6 * l, rings[l]};
    }
    // queries sorted by the same hierarchical cell key: warps then scan the same cells (less divergence, cache hits)
    const int* order = nullptr;
    if (n_query >= 4096 && n_query < (1ll << 31)) {
        uint64_t* qk_in = c.take<uint64_t>(n_query);
        uint64_t* qk_out = c.take<uint64_t>(n_query);
        int* qi_in = c.take<int>(n_query);
        int* qi_out = c.take<int>(n_query);
        size_t cub_q = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, cub_q, (uint64_t*)nullptr, (uint64_t*)nullptr, (int*)nullptr, (int*)nullptr,
                                        (int)n_query);
        void* cub_tmp_q = c.take<char>(cub_q);
        if (c.ok()) {
            const unsigned nbq = (unsigned)((n_query + T - 1) / T);
            k_knn_keys<<<nbq, T, 0, stream>>>(query_xyz, n_query, qk_in, qi_in);
            TL_LAUNCH_CHECK();
            TL_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(cub_tmp_q, cub_q, qk_in, qk_out, qi_in, qi_out, (int)n_query, 0,
                                                          3 * kFineBits, stream));
            order = qi_out;
        }
    }
    k_knn_vote<<<(unsigned)((n_query + 127) / 128), 128, 0, stream>>>(query_xyz, n_query, g, ref_labels, k, order, out_labels);
    TL_LAUNCH_CHECK();
    return TL_OK;
}

// ---- HDBSCAN ----------------------------------------------------------------------------------
size_t tl_hdbscan_workspace_bytes(int64_t n) {
    if (n <= 0) return 256;
    return grid_bytes(n) + align_up(n * 8) + align_up(n * 4) + align_up(n) + 2 * align_up(2 * 4096 * 8) + 4096;
}

int tl_core_distance(const float* points_xy, int64_t n, int32_t k, double* core, void* workspace, size_t workspace_bytes,
                     void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n == 0) return TL_OK;
    TL_REQUIRE(k >= 1 && k <= kCoreMaxK && k <= n, "tl_core_distance: k=%d must be in [1, min(%d, n=%lld)]", k, kCoreMaxK,
               (long long)n);
    TL_REQUIRE(n < (1ll << 31), "tl_core_distance: n=%lld too large", (long long)n);
    Carver c(workspace, workspace_bytes);
    CellGrid g;
    TL_REQUIRE(carve_grid(c, n, g), "tl_core_distance: workspace too small");
    const double cell = 0.5;
    int rc = build_grid<2>(points_xy, n, cell, g, stream);
    if (rc != TL_OK) return rc;
    k_core_dist<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(g.keys_out, g.idx_out, g.spts, g.seg_start, g.tkeys, g.tvals,
                                                                g.cap - 1, n, k, cell, 6, core);
    TL_LAUNCH_CHECK();
    return TL_OK;
}

int tl_mst_prim(const float* points_xy, const double* core, int64_t n, int32_t* mst_src, int32_t* mst_dst, double* mst_w,
                void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n <= 1) return TL_OK;
    TL_REQUIRE(n < (1ll << 31), "tl_mst_prim: n=%lld too large", (long long)n);
    int dev = 0, sms = 0, coop = 0;
    TL_CUDA_CHECK(cudaGetDevice(&dev));
    TL_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    TL_CUDA_CHECK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    TL_REQUIRE(coop, "tl_mst_prim: device does not support cooperative launches");
    int grid = sms;                                   // one 512-thread CTA per SM: co-resident by construction
    if ((int64_t)grid * 512 > n) grid = (int)((n + 511) / 512);
    Carver c(workspace, workspace_bytes);
    double* min_reach = c.take<double>(n);
    int* source = c.take<int>(n);
    unsigned char* in_tree = c.take<unsigned char>(n);
    double* blk_v = c.take<double>(2 * 4096);
    int* blk_j = c.take<int>(2 * 4096);
    unsigned* counter = c.take<unsigned>(64);
    TL_REQUIRE(c.ok() && grid <= 4096, "tl_mst_prim: workspace too small");
    // min_reachability = +inf (0x7ff0...), current_sources = 1, in_tree = 0 (sklearn _linkage.pyx initial state)
    TL_CUDA_CHECK(cudaMemsetAsync(in_tree, 0, n, stream));
    TL_CUDA_CHECK(cudaMemsetAsync(counter, 0, 64 * sizeof(unsigned), stream));
    k_fill_f64<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(min_reach, __longlong_as_double_host(0x7ff0000000000000ll), n);
    TL_LAUNCH_CHECK();
    k_fill_i32b<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(source, 1, n);
    TL_LAUNCH_CHECK();
    int n32 = (int)n;
    void* args[] = {(void*)&points_xy, (void*)&core, (void*)&n32, (void*)&min_reach, (void*)&source, (void*)&in_tree,
                    (void*)&blk_v, (void*)&blk_j, (void*)&counter, (void*)&mst_src, (void*)&mst_dst, (void*)&mst_w};
    TL_CUDA_CHECK(cudaLaunchCooperativeKernel((void*)k_prim, dim3(grid), dim3(512), args, 0, stream));
    count_launch();
    return TL_OK;
}

// Host pass: edges (already ordered by weight: the caller applies numpy's argsort so that ties break like sklearn's
// `_process_mst`) -> single-linkage dendrogram -> condensed tree -> stability -> EOM selection -> labels (-1 = noise).
// Restates sklearn/cluster/_hdbscan/_linkage.pyx::make_single_linkage and _tree.pyx::{_condense_tree, _compute_stability,
// _get_clusters (eom, allow_single_cluster=False, epsilon=0, no max size), _do_labelling}; traversal orders are kept
// because they fix both the floating-point summation order of the stabilities and the numbering of the clusters.
int tl_hdbscan_tree_labels(const int64_t* src, const int64_t* dst, const double* w, int64_t n, int64_t min_cluster_size,
                           int64_t* labels) {
    TL_REQUIRE(n >= 2 && src && dst && w && labels && min_cluster_size >= 2, "tl_hdbscan_tree_labels: bad arguments");
    const int64_t m = n - 1, root = 2 * m;
    // ---- single linkage (scipy hierarchy format; node n + i = merge i)
    std::vector<int64_t> parent(2 * n - 1, -1), usize(2 * n - 1, 0), left(m), right(m), csize(m);
    for (int64_t i = 0; i < n; ++i) usize[i] = 1;
    int64_t next_label = n;
    auto fast_find = [&](int64_t x) {
        int64_t p = x;
        while (parent[x] != -1) x = parent[x];
        while (parent[p] != x && parent[p] != -1) {   // path compression
            const int64_t q = parent[p];
            parent[p] = x;
            p = q;
        }
        return x;
    };
    for (int64_t i = 0; i < m; ++i) {
        const int64_t a = fast_find(src[i]), b = fast_find(dst[i]);
        TL_REQUIRE(a != b, "tl_hdbscan_tree_labels: edge %lld does not join two components (input is not a spanning tree)",
                   (long long)i);
        left[i] = a, right[i] = b, csize[i] = usize[a] + usize[b];
        parent[a] = next_label, parent[b] = next_label, usize[next_label] = csize[i];
        ++next_label;
    }
    // ---- condensed tree
    struct Row {
        int64_t parent, child;
        double lambda;
        int64_t size;
    };
    std::vector<Row> rows;
    rows.reserve(2 * n);
    auto bfs = [&](int64_t start, std::vector<int64_t>& out) {   // level order, children as (left, right)
        out.clear();
        out.push_back(start);
        for (size_t h = 0; h < out.size(); ++h) {
            const int64_t x = out[h];
            if (x >= n) out.push_back(left[x - n]), out.push_back(right[x - n]);
        }
    };
    std::vector<int64_t> order, sub;
    bfs(root, order);
    std::vector<int64_t> relabel(root + 1, 0);
    std::vector<char> ignore(root + 1, 0);
    relabel[root] = n;
    int64_t next_cluster = n + 1;
    auto spill = [&](int64_t node, int64_t from, double lam) {   // every point below `from` leaves cluster `node` at lam
        bfs(from, sub);
        for (int64_t x : sub) {
            if (x < n) rows.push_back(Row{relabel[node], x, lam, 1});
            ignore[x] = 1;
        }
    };
    for (int64_t node : order) {
        if (ignore[node] || node < n) continue;
        const int64_t l = left[node - n], r = right[node - n];
        const double dist = w[node - n];
        const double lam = dist > 0.0 ? 1.0 / dist : __longlong_as_double_host(0x7ff0000000000000ll);
        const int64_t lc = l >= n ? csize[l - n] : 1, rc = r >= n ? csize[r - n] : 1;
        if (lc >= min_cluster_size && rc >= min_cluster_size) {
            relabel[l] = next_cluster++;
            rows.push_back(Row{relabel[node], relabel[l], lam, lc});
            relabel[r] = next_cluster++;
            rows.push_back(Row{relabel[node], relabel[r], lam, rc});
        } else if (lc < min_cluster_size && rc < min_cluster_size) {
            spill(node, l, lam);
            spill(node, r, lam);
        } else if (lc < min_cluster_size) {
            relabel[r] = relabel[node];
            spill(node, l, lam);
        } else {
            relabel[l] = relabel[node];
            spill(node, r, lam);
        }
    }
    // ---- stability (cluster ids n .. next_cluster-1; the root cluster n is born at lambda 0)
    const int64_t n_clusters = next_cluster - n;
    std::vector<double> birth(n_clusters, 0.0), stab(n_clusters, 0.0);
    for (const Row& rw : rows)
        if (rw.child >= n) birth[rw.child - n] = rw.lambda;
    birth[0] = 0.0;
    for (const Row& rw : rows) stab[rw.parent - n] += (rw.lambda - birth[rw.parent - n]) * (double)rw.size;
    // ---- excess of mass over the cluster tree (children of a cluster: the two rows with size > 1, in row order)
    std::vector<std::vector<int64_t>> kids(n_clusters);
    for (const Row& rw : rows)
        if (rw.size > 1) kids[rw.parent - n].push_back(rw.child - n);
    std::vector<char> is_cluster(n_clusters, 1);
    is_cluster[0] = 0;   // allow_single_cluster = False: the root is never selected
    for (int64_t c = n_clusters - 1; c >= 1; --c) {
        double subtree = 0.0;
        for (int64_t k : kids[c]) subtree += stab[k];
        if (subtree > stab[c]) {
            is_cluster[c] = 0;
            stab[c] = subtree;
        } else {   // keep c: nothing below it is a cluster
            std::vector<int64_t> q(kids[c].begin(), kids[c].end());
            for (size_t h = 0; h < q.size(); ++h) {
                is_cluster[q[h]] = 0;
                for (int64_t k : kids[q[h]]) q.push_back(k);
            }
        }
    }
    std::vector<int64_t> label_of(n_clusters, -1);
    int64_t next = 0;
    for (int64_t c = 0; c < n_clusters; ++c)
        if (is_cluster[c]) label_of[c] = next++;
    // ---- labelling: union every row whose child is not a selected cluster; a point takes the label of the top-most
    // cluster of its set (sklearn's TreeUnionFind with union by rank yields exactly that representative)
    std::vector<int64_t> up(n + n_clusters), rank(n + n_clusters, 0);
    for (int64_t i = 0; i < n + n_clusters; ++i) up[i] = i;
    auto find = [&](int64_t x) {
        int64_t r = x;
        while (up[r] != r) r = up[r];
        while (up[x] != r) {
            const int64_t q = up[x];
            up[x] = r;
            x = q;
        }
        return r;
    };
    for (const Row& rw : rows) {
        if (rw.child >= n && is_cluster[rw.child - n]) continue;
        const int64_t xr = find(rw.parent), yr = find(rw.child);
        if (rank[xr] < rank[yr]) up[xr] = yr;
        else if (rank[xr] > rank[yr]) up[yr] = xr;
        else up[yr] = xr, ++rank[xr];
    }
    for (int64_t i = 0; i < n; ++i) {
        const int64_t c = find(i);
        labels[i] = (c >= n && c != n) ? label_of[c - n] : -1;
    }
    return TL_OK;
}

}  // extern "C"
