/* treelearn_b200 -- C ABI of the B200-native (sm_100a) TreeLearn per-tile hot path.
 *
 * The reference (ecker-lab/TreeLearn) has no C ABI: its hot path is reached through the Python
 * module boundary `tree_learn.model.TreeLearn` and the un-vendored `spconv` operator library.
 * Each entry point below names the reference interface (file:line under /root/reference) it
 * replaces.  Conventions: plain device pointers + sizes, caller (torch) owns every buffer,
 * `stream` is a cudaStream_t passed as void*, return 0 = ok / negative = error (text via
 * tl_last_error()).  No hidden allocation: scratch comes from caller workspaces whose size is
 * returned by the matching *_workspace_bytes query.  Thread-safe per stream.
 */
#ifndef TREELEARN_B200_H
#define TREELEARN_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TL_OK 0
#define TL_ERR_ARG (-1)
#define TL_ERR_CUDA (-2)
#define TL_ERR_REACH_ZERO (-3) /* spconv "reach zero!!!" (tree_learn/util/pipeline.py:91-97) */
#define TL_ERR_UNSUPPORTED (-4)

#define TL_MAX_SEG 3
#define TL_TILE_ROWS 128 /* rows per rulebook tile (tile masks, padding of index tables) */

const char* tl_last_error(void);
int tl_version(void);
/* number of kernels this library launched since the last tl_reset_launch_count() (bench `gpu_launches`) */
long long tl_launch_count(void);
void tl_reset_launch_count(void);

/* ---- point -> voxel (replaces spconv PointToVoxel.generate_voxel_with_id + the mean-pool in
 *      tree_learn/model/tree_learn.py:129-167).  Voxels come out Morton-sorted per batch element.
 * coords [N,3] f32, feats [N,F] f32 (may be NULL when F==0), batch_ids [N] i64 ascending.
 * out: voxel_keys [<=N] u64, voxel_coords [<=N,4] i32 (b,x,y,z), voxel_feats [<=N,F+3] f32 in
 * channel order [feat..., x,y,z], v2p [N] i64, *num_voxels (host).  Synchronises `stream` once. */
size_t tl_voxelize_workspace_bytes(int64_t n_points);
int tl_voxelize(const float* coords, const float* feats, int32_t n_feat, const int64_t* batch_ids,
                int64_t n_points, int32_t batch_size, float voxel_size, int32_t use_coords, int32_t use_feats,
                int32_t max_points_per_voxel, uint64_t* voxel_keys, int32_t* voxel_coords, float* voxel_feats,
                int64_t* v2p, int64_t* num_voxels, void* workspace, size_t workspace_bytes, void* stream);

/* ---- strided (k=2,s=2) level map: replaces spconv generate_conv_inds for SparseConv3d(k2,s2) and
 *      its reuse by SparseInverseConv3d (tree_learn/model/blocks.py:104-110,118-123).
 * fine_keys [n_fine] sorted -> coarse_keys [<=n_fine], coarse_coords [<=n_fine,4],
 * down_index [8][n_coarse_pad]: fine row feeding coarse row q with kappa k, or -1   (n_coarse_pad = pad128(n_fine)),
 * up_index   [8][n_fine_pad]  : coarse row feeding fine row p if kappa(p)==k, else -1,
 * down_mask [n_coarse_pad/128], up_mask [n_fine_pad/128]: per-tile bitmask of offsets in use.
 * fine_shape[3] -> coarse_shape[3] = floor((S-2)/2)+1; TL_ERR_REACH_ZERO if an axis collapses.
 * *n_coarse (host).  Synchronises `stream` once. */
size_t tl_level_workspace_bytes(int64_t n_fine);
int tl_build_level(const uint64_t* fine_keys, int64_t n_fine, const int32_t* fine_shape, uint64_t* coarse_keys,
                   int32_t* coarse_coords, int32_t* down_index, uint32_t* down_mask, int32_t* up_index,
                   uint32_t* up_mask, int32_t* coarse_shape, int64_t* n_coarse, void* workspace,
                   size_t workspace_bytes, void* stream);

/* ---- submanifold 3^3 rulebook: replaces spconv generate_subm_conv_inds (first SubMConv3d per
 *      indice_key: tree_learn/model/tree_learn.py:37-39, blocks.py:57-63).
 * keys [n] sorted unique -> nbr [27][pad128(n)] i32 (row of voxel at p+delta_k or -1; k = (dx+1)*9+(dy+1)*3+(dz+1)),
 * tile_mask [pad128(n)/128].  hash workspace: tl_rulebook_workspace_bytes(n). */
size_t tl_rulebook_workspace_bytes(int64_t n);
/* Per 128-row tile of a submanifold rulebook `nbr` [27][nbr_stride] over the Morton keys `keys` [n]: the list of DISTINCT
 * neighbour rows and the rulebook rewritten as 16-bit entries into that list, laid out so that the conv kernel's
 * shared-memory reads are bank-conflict free (csrc/tl_conv_halo.cu, "class swizzle"):
 *   halo_rows [tiles][cap]     entry p - 1 = row id | cls << 28 for list position p >= 1, -1 = unused position
 *                              (cls = keys[row] & 7, the voxel's parity class; p is odd iff cls bit 2 is set)
 *   halo_cnt  [tiles]          positions in use (0 when the tile's list exceeds `cap`)
 *   halo_lidx [tiles][28][128] rows 0..26: per LANE, position | cls << 13 of the neighbour at offset k (0 = absent);
 *                              row 27: the tile row each lane holds (the rows of a tile are permuted over the lanes so that
 *                              the 8 lanes of a shared-memory phase carry 8 different classes)
 * max_cnt (one int32, zero it first) receives the largest halo_cnt; if it exceeds `cap` (<= 2048) the caller must not pass
 * the halo to tl_conv_fwd.  New in this library: the reference's spconv gathers pair by pair (no equivalent structure). */
int tl_halo_build(const int32_t* nbr, const uint64_t* keys, int64_t n, int64_t nbr_stride, int32_t cap, int32_t* halo_rows,
                  int32_t* halo_cnt, uint16_t* halo_lidx, int32_t* max_cnt, void* stream);
int tl_subm_rulebook(const uint64_t* keys, int64_t n, const int32_t* spatial_shape, int32_t* nbr,
                     uint32_t* tile_mask, void* workspace, size_t workspace_bytes, void* stream);

/* ---- sparse convolution as a segmented gather-GEMM: replaces spconv implicit-GEMM forward for
 *      SubMConv3d / SparseConv3d / SparseInverseConv3d and the 1x1 torch.mm of Custom1x1Subm3d
 *      (tree_learn/model/blocks.py:29-39,57-70,104-123), with BatchNorm(eval)+ReLU, residual add
 *      and skip-concat fused (blocks.py:72-79,140-147; tree_learn.py:42,93).
 *   acc[r,:] = sum_seg sum_k W_seg[k] . src_seg[index_seg[k][r], :]     (index NULL => identity, n_off==1)
 *   v = acc + residual[r,:]
 *   out_raw = v ; out_act1 = relu(scale1*v+shift1) ; out_act2 = relu(scale2*v+shift2)   (each optional)
 * weight layout: mode 0 (fp32 SIMT) [n_off][c_in][c_out]; modes 1/2 (tcgen05 tf32 / f16) [n_off][c_in/32][c_out][32]
 * with the 16 B chunks of every 32-channel row XOR-swizzled by the row index (c ^ (n & 7) for 128 B tf32 rows,
 * c ^ ((n >> 1) & 3) for 64 B fp16 rows) = the UMMA shared-memory B-operand image (treelearn_b200/sparse.py::pack_weight_tc). */
typedef struct {
    const float* src;
    int64_t src_stride; /* floats per row */
    int32_t c_in;
    int32_t n_off;
    const int32_t* index; /* [n_off][index_stride] or NULL */
    int64_t index_stride;
    const uint32_t* tile_mask; /* per 128-row tile, bit k = offset k has >=1 pair; NULL = all */
    const float* weight;
} tl_conv_seg;

typedef struct {
    int32_t n_out;
    int32_t c_out;
    int32_t n_seg;
    int32_t src_fp32_mask; /* modes 2/3: bit s set = segment s reads raw fp32 rows [*, c_in] (the residual stream) and
                              converts them to the operand format in registers; 0 = every source is in the operand format */
    tl_conv_seg seg[TL_MAX_SEG];
    const float* residual; /* [n_out, c_out] or NULL */
    float* out_raw;
    float* out_act1;
    const float* scale1;
    const float* shift1;
    float* out_act2;
    const float* scale2;
    const float* shift2;
    float* splitk_ws; /* optional [n_out, c_out] fp32 scratch: lets layers with few row tiles run split-K */
    /* optional (modes 2/3, every segment a 27-offset submanifold segment over the SAME rulebook): the level's halo lists
     * from tl_halo_build.  With them the conv fetches each distinct neighbour row of a 128-row tile once into shared
     * memory (csrc/tl_conv_halo.cu).  halo_umax = largest halo_cnt of the level (must be <= halo_cap). */
    const int32_t* halo_rows;  /* [tiles][halo_cap] */
    const int32_t* halo_cnt;   /* [tiles] */
    const uint16_t* halo_lidx; /* [tiles][28][128], see tl_halo_build */
    int32_t halo_cap;
    int32_t halo_umax;
} tl_conv_desc;

#define TL_MODE_FP32 0
#define TL_MODE_TF32 1
#define TL_MODE_F16 2 /* tcgen05 kind::f16: segment sources, weights and the activated outputs (out_act1/2) are fp16;
                         residual, out_raw and the accumulation stay fp32 */
#define TL_MODE_F16X2 3 /* two-term fp16 split of both operands (x = hi + lo, three tcgen05.mma per K step, fp32 accumulate):
                           products carry ~22 mantissa bits = the reference's fp32 arithmetic (tree_learn inference runs spconv
                           in fp32: configs/pipeline/pipeline.yaml:12 `fp16` is never read).  Segment sources and activated
                           outputs are [row][C/32][2][32] fp16 (per 32-channel block: 64 B of hi halves, 64 B of lo halves);
                           weights [n_off][c_in/32][2][c_out][32] fp16.
                           The tensor-memory-A kernel behind modes 2 and 3 (csrc/tl_conv_ts.cu) keeps every tensor it reads or
                           writes -- fp32 residual / raw output, activated operands -- in "P-layout": inside each 32-channel
                           block, position 8q + 2g + e holds logical channel 8g + 2q + e (q, g = 0..3, e = 0..1); the weights
                           carry the matching K order (treelearn_b200/sparse.py: PI, to_p / from_p, pack_weight_ts).  scale /
                           shift stay in logical order.  TL_TS=0 selects round 1's shared-memory-A kernel (natural layout). */
int tl_conv_fwd(const tl_conv_desc* desc, int32_t mode, void* stream);

/* ---- voxel -> point gather + the two MLP heads: replaces `features[v2p_map]` and MLP forward
 *      (tree_learn/model/tree_learn.py:97-103, blocks.py:8-18).  BN(eval) is folded into w1/b1 by the host.
 * voxel_feats [M,C]; v2p [N]; per head h in {sem(2), off(3)}: w1 [C][C] (row = out), b1 [C], w2 [O][C], b2 [O]. */
int tl_heads_fwd(const void* voxel_feats, int32_t feats_half /* 0: fp32 rows; 1: fp16 rows; 2: the TL_MODE_F16X2 operand format
                 [M][C/32][2][32] fp16 in P-layout; 3: fp16 rows in P-layout (TL_MODE_F16 backbone, tensor-memory-A kernel) */,
                 const int64_t* v2p, int64_t n_points, int32_t channels,
                 const float* sem_w1, const float* sem_b1, const float* sem_w2, const float* sem_b2,
                 const float* off_w1, const float* off_b1, const float* off_w2, const float* off_b2,
                 float* backbone_feats, float* sem_logits, float* offsets, void* stream);

/* ---- overlap merge: replaces `ensemble` (tree_learn/util/pipeline.py:113-141): group rows by
 *      round(coords,2) and average `n_val` float columns; output sorted by (x,y,z).
 * values [n, n_val] f32 -> out_coords [<=n,3], out_values [<=n,n_val], *n_groups (host). */
size_t tl_merge_workspace_bytes(int64_t n);
int tl_merge_groupby_mean(const float* coords, const float* values, int64_t n, int32_t n_val, float* out_coords,
                          float* out_values, int32_t* group_of_row, int64_t* n_groups, void* workspace,
                          size_t workspace_bytes, void* stream);

/* ---- DBSCAN(eps, min_samples=2)-equivalent clustering + size filter + consecutive relabel:
 *      replaces `group_dbscan` (tree_learn/util/pipeline.py:173-180, 195-206).
 * points [n,2] f32 -> labels [n] i64: start_num.. for clusters with >= min_cluster_size points
 * (numbered by lowest member index), else not_assigned.  *n_clusters (host). */
size_t tl_cluster_workspace_bytes(int64_t n);
int tl_cluster_radius_cc(const float* points_xy, int64_t n, double radius, int64_t min_cluster_size,
                         int64_t not_assigned_label, int64_t start_num, int64_t* labels, int64_t* n_clusters,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ---- kNN(k) majority vote: replaces `assign_remaining_points_nearest_neighbor`
 *      (tree_learn/util/pipeline.py:287-296).  ref [nr,3] f32 + ref_labels [nr] i64; query [nq,3] ->
 * out_labels [nq] (most frequent label among the k nearest, ties -> smallest label). */
size_t tl_knn_workspace_bytes(int64_t n_ref, int64_t n_query);
int tl_knn_vote(const float* ref_xyz, const int64_t* ref_labels, int64_t n_ref, const float* query_xyz,
                int64_t n_query, int32_t k, int64_t* out_labels, void* workspace, size_t workspace_bytes,
                void* stream);

/* ---- HDBSCAN(min_cluster_size = m) on 2-D points: replaces `group_hdbscan` -> sklearn.cluster.HDBSCAN
 *      (tree_learn/util/pipeline.py:184-191; the DEFAULT clusterer, configs/_modular/grouping.yaml:7).
 * tl_core_distance: core[i] = distance from point i to its k-th nearest neighbour, itself included (fp64, no FMA) --
 *                   sklearn `_hdbscan_prims`: NearestNeighbors(n_neighbors=min_samples).kneighbors(X)[:, -1].
 * tl_mst_prim:      minimum spanning tree of the mutual-reachability graph max(core_a, core_b, |a-b|), Prim's algorithm
 *                   from node 0 with sklearn's tie rules (`mst_from_data_matrix`): edge i = (mst_src[i], mst_dst[i], mst_w[i])
 *                   in the order Prim adds them.  One persistent cooperative kernel.
 * tl_hdbscan_tree_labels: HOST arrays.  Edges sorted by weight (caller sorts) -> single linkage -> condensed tree ->
 *                   stability -> excess-of-mass selection -> labels[n] (0.. in sklearn's numbering, -1 = noise). */
size_t tl_hdbscan_workspace_bytes(int64_t n);
int tl_core_distance(const float* points_xy, int64_t n, int32_t k, double* core, void* workspace, size_t workspace_bytes,
                     void* stream);
int tl_mst_prim(const float* points_xy, const double* core, int64_t n, int32_t* mst_src, int32_t* mst_dst, double* mst_w,
                void* workspace, size_t workspace_bytes, void* stream);
int tl_hdbscan_tree_labels(const int64_t* src, const int64_t* dst, const double* w, int64_t n, int64_t min_cluster_size,
                           int64_t* labels);

/* ---- training (SURVEY §8 a12 train mode, a16): BatchNorm1d(eps, momentum) with batch statistics applied to the
 *      feature rows of a sparse tensor + ReLU (tree_learn/model/tree_learn.py:34, blocks.py:57-70) and its backward,
 *      and the sparse-conv weight gradient (autograd through spconv in tools/training/train.py:40).
 *      The data gradient of a sparse conv is tl_conv_fwd with transposed weights on the transposed rulebook
 *      (3^3 submanifold table: same table, offsets mirrored k -> 26-k; strided maps: down_index <-> up_index).
 * tl_bn_stats:    acc[0..c) = sum_r x[r,j], acc[c..2c) = sum_r x[r,j]^2        (fp64, zeroed by the call)
 * tl_bn_finalize: mean, invstd = 1/sqrt(biased var + eps), scale = gamma*invstd, shift = beta - mean*scale;
 *                 running_mean/var (nullable) get the momentum update (unbiased variance), like torch.
 * tl_bn_relu_apply: out = relu(scale*x + shift)
 * tl_bn_relu_bwd: dy = d_act * [scale*x+shift > 0]; dbeta = sum dy; dgamma = sum dy*xhat;
 *                 batch_stats=1: dx = scale*(dy - dbeta/n - xhat*dgamma/n); batch_stats=0 (eval/frozen): dx = scale*dy.
 *                 acc: [2c] fp64 scratch. */
int tl_bn_stats(const float* x, int64_t n, int32_t c, double* acc, void* stream);
int tl_bn_finalize(const double* acc, int64_t n, int32_t c, const float* gamma, const float* beta, float eps,
                   float momentum, float* running_mean, float* running_var, float* mean, float* invstd, float* scale,
                   float* shift, void* stream);
int tl_bn_relu_apply(const float* x, int64_t n, int32_t c, const float* scale, const float* shift, float* out,
                     void* stream);
int tl_bn_relu_bwd(const float* x, const float* d_act, int64_t n, int32_t c, const float* scale, const float* shift,
                   const float* mean, const float* invstd, int32_t batch_stats, double* acc, float* dx, float* dgamma,
                   float* dbeta, void* stream);
/* Pack a conv parameter (spconv KRSC layout [C_out][n_off][C_in], fp32) into the tcgen05 B-operand image that tl_conv_fwd
 * modes 1/2 read: forward layout (transpose = 0) or the data-gradient layout (transpose = 1: C_in/C_out swapped, offsets
 * mirrored when `mirror`); fp16 (`half`) or TF32-rounded fp32; bk = channels per chunk (32, or 64 for fp16 with C_in' % 64 == 0). */
int tl_pack_weight_tc(const float* w, int32_t c_out, int32_t n_off, int32_t c_in, int32_t transpose, int32_t mirror,
                      int32_t half, int32_t bk, void* out, void* stream);
/* dw[k][ci][co] = sum_r src[index[k][r], ci] * d_out[r, co]   (index NULL => identity, n_off == 1); dw is overwritten.
 * tf32 = 0: fp32 FMA (exact-arithmetic training mode); tf32 = 1: TF32 tensor-core products, fp32 accumulation. */
int tl_conv_wgrad(const float* src, int64_t src_stride, int32_t c_in, int32_t n_off, const int32_t* index,
                  int64_t index_stride, const uint32_t* tile_mask, const float* d_out, int64_t n_out, int32_t c_out,
                  float* dw, int32_t tf32, void* stream);
/* The same weight gradient on the tcgen05 tensor cores (TF32 operands = the upper 19 bits of the fp32 inputs, fp32
 * accumulate in tensor memory; csrc/tl_wgrad_tc.cu) for C_in % 32 == 0, C_out % 32 == 0, C_out <= 256
 * (tl_conv_wgrad_tc_eligible).  Replaces spconv's implicit-GEMM wgrad reached through autograd at
 * tools/training/train.py:40.  Per-CTA partial sums in `workspace` are added in a fixed order: run-to-run deterministic. */
int tl_conv_wgrad_tc_eligible(int32_t c_in, int32_t c_out);
size_t tl_conv_wgrad_tc_workspace_bytes(int64_t n_out, int32_t c_in, int32_t n_off, int32_t c_out);
int tl_conv_wgrad_tc(const float* src, int64_t src_stride, int32_t c_in, int32_t n_off, const int32_t* index,
                     int64_t index_stride, const uint32_t* tile_mask, const float* d_out, int64_t n_out, int32_t c_out,
                     float* dw, void* workspace, size_t workspace_bytes, void* stream);

/* ---- after the path -----------------------------------------------------------
 * tl_hash_join_last: exact-coordinate join, replaces the Python `hash(tuple(point))` dictionaries of
 *      `propagate_preds_hash_vox` (tree_learn/util/pipeline.py:455-465).  Rows are [n,3] fp32 (`*_f64` = 0) or fp64 (1);
 *      `*_round2` = 1 rounds every coordinate to two decimals first, exactly as numpy.round(x, 2) does in the array's
 *      dtype (multiply by 100, rint, divide).  Keys compare as fp64 values (-0.0 == 0.0).  out_vals[j] = build_vals of
 *      the LAST build row whose key equals probe row j (dict(zip(...)) semantics), or `missing`. */
size_t tl_hash_join_workspace_bytes(int64_t n_build);
int tl_hash_join_last(const void* build_xyz, int32_t build_f64, int32_t build_round2, const int64_t* build_vals,
                      int64_t n_build, const void* probe_xyz, int32_t probe_f64, int32_t probe_round2, int64_t n_probe,
                      int64_t missing, int64_t* out_vals, void* workspace, size_t workspace_bytes, void* stream);
/* tl_cooccurrence_counts: the point counts behind `get_detections`' IoU / precision / recall matrices
 *      (tree_learn/util/eval.py:7-31, get_eval_components :230-238).  counts is [(n_pred+1) x (n_gt+1)] u64, overwritten:
 *      counts[p][g] = #points with pred == p and gt == g; labels outside [0, n) fall into the last row / column, so
 *      row p sums to |pred == p| and column g to |gt == g|. */
int tl_cooccurrence_counts(const int64_t* pred, const int64_t* gt, int64_t n, int64_t n_pred, int64_t n_gt,
                           unsigned long long* counts, void* stream);

/* ---- before the path ------------------------------------------------------------
 * tl_voxel_downsample_trace: replaces `voxelize` -> open3d 0.17 PointCloud.voxel_down_sample_and_trace
 *      (tree_learn/util/data_preparation.py:60-79).  points [n,3] fp64 (`round2_first` = numpy.round(points, 2) on the fly,
 *      :62); voxel index = floor((p - voxel_min_bound) / voxel_size) in fp64 (open3d: voxel_min_bound = min_bound -
 *      voxel_size / 2).  Per voxel, in ascending (ix, iy, iz) order (open3d's order is std::unordered_map's):
 *      out_points [m,3] = fp64 sum in input order / count, first_index [m] = lowest input index (the row whose extra
 *      columns `voxelize` keeps), offsets [m+1] + trace [n] = CSR of the input indices per voxel, in input order.
 *      out_points / first_index need room for n rows, offsets for n+1.  *n_voxels (host).  Synchronises `stream`. */
size_t tl_downsample_workspace_bytes(int64_t n_points);
int tl_voxel_downsample_trace(const double* points, int64_t n_points, int32_t round2_first, double voxel_size,
                              double voxel_min_bound, double* out_points, int64_t* first_index, int64_t* offsets,
                              int64_t* trace, int64_t* n_voxels, void* workspace, size_t workspace_bytes, void* stream);
/* tl_verticality: replaces `compute_features(..., feature_names=['verticality'])` -> jakteristics 0.5.1
 *      (tree_learn/util/data_preparation.py:83-88).  points [n,3] fp64 -> out [n] fp64: 1 - |n_z| with n the unit
 *      eigenvector of the smallest eigenvalue of the covariance of all points within `search_radius` (closed ball, the
 *      point itself included); NaN when fewer than 3 points are in the ball.  Synchronises `stream` once. */
size_t tl_verticality_workspace_bytes(int64_t n_points);
int tl_verticality(const double* points, int64_t n_points, double search_radius, double* out, void* workspace,
                   size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif
