"""Pre-path row: voxel down-sampling with trace, verticality, tile cutting.  Tile cutting is pinned to
golden vectors recorded from the reference's own SampleGenerator (tests/golden/make_golden_tiles.py); the down-sampler
and the verticality kernel are checked against the oracle's restatement of open3d 0.17 / jakteristics 0.5.1 (both absent
from this image: parity unpinned)."""
import os

import numpy as np
import pytest
import torch

from oracle import prepare_ref
from treelearn_b200 import prepare, synth

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
CASES = ['a', 'b', 'c']


@pytest.fixture(scope='module')
def gold():
    return np.load(os.path.join(GOLD, 'tiles_small.npz'))


def _check_tiles_against_golden(g, name, tiles):
    assert len(tiles) == int(g[f'{name}:n_tiles'])
    assert np.array_equal(np.array([len(t['points']) for t in tiles]), g[f'{name}:rows'])
    assert np.array_equal(np.array([t['center'] for t in tiles]), g[f'{name}:centers'])
    assert np.array_equal(np.array([t['points'].astype(np.float64).sum(axis=0) for t in tiles]), g[f'{name}:sum_points'])
    assert np.array_equal(np.array([t['instance_label'].astype(np.int64).sum() for t in tiles]), g[f'{name}:sum_labels'])
    for t in (0, len(tiles) - 1):
        for key in ('points', 'feat', 'instance_label'):
            want = g[f'{name}:tile{t}:{key}']
            assert tiles[t][key].dtype == want.dtype and np.array_equal(tiles[t][key], want), (t, key)


def _raw_plot(seed, n=30000):
    """An un-voxelised cloud: several points per 0.1 m voxel, coordinates with more than 2 decimals, a label column."""
    f = synth.synth_forest(edge=8.0, height=6.0, n_trees=3, seed=seed, ground_density=60.0)
    rng = np.random.default_rng(seed)
    pick = rng.integers(0, len(f['coords']), n)
    pts = f['coords'][pick].astype(np.float64) + rng.normal(0, 0.03, (n, 3)) + np.array([431.2, -77.7, 12.0])
    return np.hstack([pts, f['inst'][pick].astype(np.float64).reshape(-1, 1)])


# ---- CPU ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('name', CASES)
def test_oracle_tiles_match_reference_golden(gold, name):
    inner_edge, outer_edge, stride = gold[f'{name}:cfg']
    tiles = prepare_ref.cut_tiles_ref(gold[f'{name}:points'], gold[f'{name}:labels'], gold[f'{name}:feats'],
                                      int(inner_edge), float(outer_edge), float(stride))
    _check_tiles_against_golden(gold, name, tiles)


@pytest.mark.parametrize('name', CASES)
def test_tile_grid_and_cutting_host_logic_match_reference_golden(gold, name):
    """The grid arithmetic and the torch masking are device-agnostic host logic: run here on CPU tensors."""
    inner_edge, outer_edge, stride = gold[f'{name}:cfg']
    tiles = prepare.cut_tiles(gold[f'{name}:points'], gold[f'{name}:labels'], gold[f'{name}:feats'], int(inner_edge),
                              float(outer_edge), float(stride), device='cpu')
    _check_tiles_against_golden(gold, name, tiles)


def test_oracle_downsample_properties():
    data = _raw_plot(5, n=4000)
    down, trace = prepare_ref.voxelize_ref(data, 0.1)
    assert sorted(np.concatenate(trace).tolist()) == list(range(len(data)))           # a partition of the input
    assert all(np.all(np.diff(t) > 0) for t in trace)                                  # input order inside a voxel
    pts = np.round(data[:, :3], 2)
    assert all(np.ptp(pts[t], axis=0).max() <= 0.1 + 1e-9 for t in trace)               # one voxel = one 0.1 m cube
    assert np.allclose(down[7, :3], pts[trace[7]].mean(axis=0)) and down[7, 3] == data[trace[7][0], 3]


def test_trace_behaves_like_a_list_of_index_vectors():
    tr = prepare.Trace(np.array([0, 2, 3, 6]), np.array([4, 9, 1, 0, 5, 7]))
    assert len(tr) == 3 and [list(item) for item in tr] == [[4, 9], [1], [0, 5, 7]]
    assert [item[0] for item in tr] == [4, 1, 0] and list(tr[-1]) == [0, 5, 7]
    with pytest.raises(IndexError):
        tr[3]


def test_prepare_functions_need_the_gpu():
    if torch.cuda.is_available():
        pytest.skip('CPU-only check')
    from treelearn_b200._lib import TreeLearnCudaError
    with pytest.raises(TreeLearnCudaError):
        prepare.voxelize(np.zeros((4, 4)), 0.1)
    with pytest.raises(TreeLearnCudaError):
        prepare.compute_features(np.zeros((4, 3)))
    with pytest.raises(TreeLearnCudaError):
        prepare.cut_tiles(np.zeros((4, 3), np.float32), np.zeros(4, np.float32), np.zeros((4, 1), np.float32), 8, 13.5, 0.5)


# ---- GPU ------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize('seed', [5, 6])
def test_voxelize_cuda_bit_exact_vs_oracle(seed):
    data = _raw_plot(seed)
    down, trace = prepare.voxelize(data, 0.1)
    want, want_trace = prepare_ref.voxelize_ref(data, 0.1)
    assert down.dtype == np.float64 and np.array_equal(down, want)
    assert len(trace) == len(want_trace)
    assert np.array_equal(trace.indices, np.concatenate(want_trace))
    assert np.array_equal(np.diff(trace.offsets), np.array([len(t) for t in want_trace]))
    assert len(want) < len(data) / 2                                   # the fixture really has several points per voxel


@pytest.mark.gpu
def test_voxelize_cuda_edge_cases():
    down, trace = prepare.voxelize(np.zeros((0, 4)), 0.1)
    assert down.shape == (0, 4) and len(trace) == 0
    one = np.array([[1.234, 5.678, 9.0, 3.0]])
    down, trace = prepare.voxelize(one, 0.1)
    assert np.array_equal(down, np.array([[1.23, 5.68, 9.0, 3.0]])) and [list(t) for t in trace] == [[0]]
    xyz_only = np.array([[0.01, 0.01, 0.01], [0.02, 0.03, 0.04], [5.0, 5.0, 5.0]])
    down, trace = prepare.voxelize(xyz_only, 0.1)
    want, want_trace = prepare_ref.voxelize_ref(xyz_only, 0.1)
    assert np.array_equal(down, want) and [list(t) for t in trace] == [list(t) for t in want_trace]
    with pytest.raises(Exception, match='voxel_size'):
        prepare.voxelize(one, 0.0)


@pytest.mark.gpu
def test_verticality_cuda_vs_oracle():
    f = synth.synth_forest(edge=7.0, height=7.0, n_trees=3, seed=9, ground_density=90.0)
    # un-rounded coordinates: no pair sits exactly on the search radius, where a kd-tree's box pruning and a direct
    # distance test may legitimately disagree
    pts = f['coords'].astype(np.float64) + np.random.default_rng(1).normal(0, 0.004, f['coords'].shape)
    pts = np.vstack([pts, [[50.0, 50.0, 50.0]], [[60.0, 60.0, 60.0], [60.1, 60.0, 60.0]]])    # 1 and 2 points in the ball
    want, gap = prepare_ref.verticality_ref(pts, 0.6, return_gap=True)
    got = prepare.verticality_cuda(torch.from_numpy(pts).cuda(), 0.6).cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(want)) and np.isnan(want[-3:]).all()
    ok = ~np.isnan(want) & (gap > 1e-6)
    assert ok.mean() > 0.95
    assert np.abs(got[ok] - want[ok]).max() < 1e-7                 # fp64 both sides; tolerance = conditioning of the normal
    feats = prepare.compute_features(pts, 0.6)
    ref = prepare_ref.compute_features_ref(pts, 0.6)
    assert feats.dtype == np.float32 and feats.shape == (len(pts), 1) and not np.isnan(feats).any()
    assert np.abs(feats[ok] - ref[ok]).max() < 1e-6
    assert abs(feats[-1, 0] - got[~np.isnan(got)].mean()) < 1e-6 and abs(feats[-1, 0] - ref[-1, 0]) < 2e-2    # NaN -> column mean


@pytest.mark.gpu
@pytest.mark.parametrize('name', CASES)
def test_cut_tiles_cuda_matches_reference_golden(gold, name):
    inner_edge, outer_edge, stride = gold[f'{name}:cfg']
    tiles = prepare.cut_tiles(gold[f'{name}:points'], gold[f'{name}:labels'], gold[f'{name}:feats'], int(inner_edge),
                              float(outer_edge), float(stride))
    _check_tiles_against_golden(gold, name, tiles)


@pytest.mark.gpu
def test_prepare_tiles_end_to_end_in_memory():
    data = _raw_plot(8, n=60000)
    plot, feats, tiles, trace = prepare.prepare_tiles(data, 0.1, 0.6, inner_edge=4, outer_edge=3.0, stride=0.5)
    assert plot.dtype == np.float32 and plot.shape[1] == 4 and feats.shape == (len(plot), 1) and len(trace) == len(plot)
    assert np.array_equal(plot, np.round(plot, 2)) and 0.0 <= feats.min() and feats.max() <= 1.0
    want = prepare_ref.cut_tiles_ref(plot[:, :3], plot[:, 3], feats, 4, 3.0, 0.5)
    assert len(tiles) == len(want) > 1
    for a, b in zip(tiles, want):
        for key in ('points', 'feat', 'instance_label', 'center'):
            assert np.array_equal(a[key], b[key]), key
    inner = sum(int((np.abs(t['points'][:, :2]).max(axis=1) <= 2.0 + 1e-3).sum()) for t in tiles)
    assert inner >= len(plot)                                   # the inner squares (overlapping at stride 0.5) cover the plot


# ---- tiles -> model input (treelearn_b200/plot.py), CPU -----------------------------------------------------------------
@pytest.mark.parametrize('name', ['b', 'c'])
def test_tiles_to_batches_match_reference_dataset_golden(gold, name):
    """`tiles_to_batches` against the batch the reference's TreeDataset (test mode) + collate_fn built from the first two
    golden tiles: same keys, dtypes and values (offset labels, masks, centres, batch ids)."""
    from treelearn_b200 import plot
    inner_edge, outer_edge, stride = gold[f'{name}:cfg']
    tiles = prepare_ref.cut_tiles_ref(gold[f'{name}:points'], gold[f'{name}:labels'], gold[f'{name}:feats'],
                                      int(inner_edge), float(outer_edge), float(stride))
    batch = next(plot.tiles_to_batches(tiles[:2], int(inner_edge), batch_size=2))
    keys = [k.split(':')[-1] for k in gold.files if k.startswith(f'{name}:batch:')]
    assert sorted(keys) == sorted(batch.keys())
    for key in keys:
        want = gold[f'{name}:batch:{key}']
        got = batch[key].numpy() if torch.is_tensor(batch[key]) else np.asarray(batch[key])
        assert got.dtype == want.dtype and np.array_equal(got, want), key
    assert batch['masks_inner'].any() and not batch['masks_inner'].all()
    assert bool(batch['masks_off'].any()) == (name == 'c')          # case c has a tree inside the first inner squares
