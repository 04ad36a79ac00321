"""ORACLE (test infrastructure): CPU restatement of the step before the per-tile path:
voxel down-sampling with trace, the verticality feature and tile cutting.

  voxel_down_sample_and_trace_ref   tree_learn/util/data_preparation.py:60-79 -> open3d 0.17.0 (setup/requirements.txt:9)
        `PointCloud::VoxelDownSampleAndTrace`: voxel index = floor((p - (min_bound - voxel/2)) / voxel) in fp64, the
        voxel's point = fp64 sum of its points in input order / count, trace = the input indices in input order.
        open3d emits voxels in std::unordered_map order (implementation-defined); here: ascending (ix, iy, iz).
        **open3d is absent from this image: parity unpinned** (published algorithm restated, no golden vectors).
  verticality_ref                   data_preparation.py:83-88 -> jakteristics 0.5.1 (requirements.txt:10)
        `compute_features(points, search_radius, feature_names=['verticality'])`: neighbours = cKDTree.query_ball_point
        (closed ball, the point itself included), covariance of the neighbours, eigenvectors sorted by decreasing
        eigenvalue, verticality = 1 - |z component of the third one|; fewer than 3 neighbours -> NaN.
        **jakteristics is absent from this image: parity unpinned.**
  replace_nanfeatures_ref           data_preparation.py:91-100
  tile_grid_ref / cut_tiles_ref     data_preparation.py:333-494 (`SampleGenerator.tile_generate_and_save`, plot_corners=None,
        no denoising = the configured default, configs/_modular/sample_generation.yaml:9-14), same scalar types per
        operation.  Pinned against the reference's OWN class by tests/golden/make_golden_tiles.py -> tiles_small.npz.
numpy + scipy only.  Only tests/, smoke() and bench.py's cpu_baseline legs may import it.
"""
import numpy as np
from scipy.spatial import cKDTree


def voxel_down_sample_and_trace_ref(points, voxel_size):
    """points [n,3] (rounded to 2 decimals first, data_preparation.py:62) -> (voxel points [m,3] f64, list of index arrays)."""
    pts = np.round(np.asarray(points, dtype=np.float64), 2)
    bound = np.max(np.abs(pts)) + 100
    voxel_min = -bound - voxel_size * 0.5
    cells = {}
    for i, p in enumerate(pts):
        key = tuple(int(np.floor(v)) for v in (p - voxel_min) / voxel_size)
        acc = cells.setdefault(key, [np.zeros(3), []])
        acc[0] += p
        acc[1].append(i)
    keys = sorted(cells)
    down = np.array([cells[k][0] / float(len(cells[k][1])) for k in keys]).reshape(-1, 3)
    return down, [np.array(cells[k][1], dtype=np.int64) for k in keys]


def voxelize_ref(data, voxel_size):
    """`voxelize` (data_preparation.py:60-79): extra columns come from the first point of every voxel."""
    data = np.asarray(data)
    down, trace = voxel_down_sample_and_trace_ref(data[:, :3], voxel_size)
    if data.shape[1] >= 4:
        down = np.hstack((down, data[:, 3:][[t[0] for t in trace]]))
    return down, trace


def verticality_ref(points, search_radius, return_gap=False):
    """[n] verticality; with return_gap also (lambda_1 - lambda_0) / lambda_2 of the ascending eigenvalues: where it is
    ~0 the normal direction is not defined (collinear / isotropic neighbourhoods) and implementations may differ."""
    pts = np.asarray(points, dtype=np.float64)
    tree = cKDTree(pts)
    out, gap = np.full(len(pts), np.nan), np.zeros(len(pts))
    for i, nb in enumerate(tree.query_ball_point(pts, search_radius)):
        if len(nb) < 3:
            continue
        w, v = np.linalg.eigh(np.cov(pts[nb].T))           # ascending eigenvalues: column 0 = the surface normal
        out[i] = 1.0 - abs(v[2, 0])
        gap[i] = (w[1] - w[0]) / w[2] if w[2] > 0 else 0.0
    return (out, gap) if return_gap else out


def replace_nanfeatures_ref(features):
    features = np.array(features, dtype=np.float64, copy=True)
    mean = np.nanmean(features, axis=0)
    for c in range(features.shape[1]):
        features[np.isnan(features[:, c]), c] = mean[c]
    return features


def compute_features_ref(points, search_radius=0.6):
    return replace_nanfeatures_ref(verticality_ref(points, search_radius)[:, None]).astype(np.float32)


def tile_grid_ref(x_range, y_range, inner_edge, outer_edge, stride):
    """Inner squares [T,4] = (xmin, xmax, ymin, ymax) f64 in the reference's row-major order (rows from the top)."""
    xmin = np.round(x_range[0] - 1.5 * outer_edge, 2)
    xmax = np.round(x_range[1] + 1.5 * outer_edge, 2)
    ymin = np.round(y_range[0] - 1.5 * outer_edge, 2)
    ymax = np.round(y_range[1] + 1.5 * outer_edge, 2)
    ncols = int(np.round((xmax - xmin - 2 * outer_edge) / inner_edge))
    edge_x = np.round((xmax - xmin - 2 * outer_edge) / ncols, 5)
    ncols = int((ncols - 1) / stride + 1)
    nrows = int(np.round((ymax - ymin - 2 * outer_edge) / inner_edge))
    edge_y = np.round((ymax - ymin - 2 * outer_edge) / nrows, 5)
    nrows = int((nrows - 1) / stride + 1)
    inner = np.empty((nrows * ncols, 4))
    for i in range(nrows):
        for j in range(ncols):
            inner[i * ncols + j] = [xmin + outer_edge + stride * j * edge_x, xmin + outer_edge + (stride * j + 1) * edge_x,
                                    ymax - outer_edge - (stride * i + 1) * edge_y, ymax - outer_edge - stride * i * edge_y]
    return np.round(inner, 5)


def cut_tiles_ref(points, labels, feats, inner_edge, outer_edge, stride):
    """points [n,3] f32, labels [n] f32, feats [n,F] f32 (what SampleGenerator.__init__ loads) -> list of tile dicts
    with the keys the reference saves (data_preparation.py:472-476)."""
    points = np.asarray(points)
    rows = np.hstack([np.hstack((points, np.asarray(labels).reshape(-1, 1))), feats])
    x, y = rows[:, 0], rows[:, 1]
    inner = tile_grid_ref((x.min(), x.max()), (y.min(), y.max()), inner_edge, outer_edge, stride)
    outer = inner + np.array([-outer_edge, outer_edge, -outer_edge, outer_edge]).reshape(1, 4)
    tiles = []
    for sq_in, sq_out in zip(inner, outer):
        lo_x, hi_x, lo_y, hi_y = (np.float32(v) for v in sq_out)    # torch compares an fp32 tensor with a 0-dim fp64 in fp32
        chunk = rows[(x >= lo_x) & (x <= hi_x) & (y >= lo_y) & (y <= hi_y)]
        cx, cy = chunk[:, 0], chunk[:, 1]
        if not ((cx >= sq_in[0]) & (cx < sq_in[1]) & (cy > sq_in[2]) & (cy <= sq_in[3])).any():
            continue
        sq32 = sq_in.astype(np.float32)
        center_x, center_y = np.round((sq32[0] + sq32[1]) / 2, 6), np.round((sq32[2] + sq32[3]) / 2, 6)
        shift = np.concatenate([np.array([center_x, center_y, 0, 0]), np.zeros(rows.shape[1] - 4)]).reshape(1, -1)
        chunk = (chunk.astype(np.float64) - shift).astype(np.float32)
        tiles.append({'points': chunk[:, :3], 'feat': chunk[:, 4:], 'instance_label': chunk[:, 3].astype(np.int32),
                      'center': np.array([center_x, center_y, 0])})
    return tiles
