"""Post-path rows: `propagate_preds`, `propagate_preds_hash_vox`, `get_detections` -- oracle and CUDA
path against golden vectors recorded from the reference's own functions (tests/golden/make_golden_post.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import post_ref
from treelearn_b200 import post

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
KNN, VOX, DET = ['knn_a', 'knn_b'], ['vox_f64', 'vox_f32'], ['det_a', 'det_b']
DET_KEYS = ['matched_gts', 'matched_preds', 'iou', 'prec', 'rec']


@pytest.fixture(scope='module')
def gold():
    return np.load(os.path.join(GOLD, 'post_small.npz'))


def _counts_numpy(pred, gt, n_pred, n_gt):
    p = np.where((pred < 0) | (pred >= n_pred), n_pred, pred)
    g = np.where((gt < 0) | (gt >= n_gt), n_gt, gt)
    return np.bincount(p * (n_gt + 1) + g, minlength=(n_pred + 1) * (n_gt + 1)).reshape(n_pred + 1, n_gt + 1)


# ---- CPU ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('name', KNN)
def test_oracle_propagate_preds_matches_reference_golden(gold, name):
    got = post_ref.propagate_preds_ref(gold[f'{name}:src'], gold[f'{name}:lab'], gold[f'{name}:tgt'], 5)
    assert np.array_equal(got, gold[f'{name}:out'])


@pytest.mark.parametrize('name', VOX)
def test_oracle_hash_vox_matches_reference_golden(gold, name):
    got, missing = post_ref.propagate_preds_hash_vox_ref(gold[f'{name}:cur'], gold[f'{name}:preds'], gold[f'{name}:ret'])
    assert np.array_equal(got, gold[f'{name}:out']) and np.array_equal(missing, gold[f'{name}:missing'])
    assert missing.any() and not missing.all()


def test_oracle_round2_is_numpy_round():
    rng = np.random.default_rng(0)
    for dtype in (np.float32, np.float64):
        a = np.concatenate([rng.uniform(-500, 500, 20000), np.arange(-200, 200) / 200.0 + 0.005]).astype(dtype)
        assert np.array_equal(post_ref.np_round2(a), np.round(a, 2)) and post_ref.np_round2(a).dtype == dtype


@pytest.mark.parametrize('name', DET)
def test_oracle_and_host_detection_matrices_match_reference_golden(gold, name):
    gt, pred, non_tree = gold[f'{name}:gt'], gold[f'{name}:pred'], int(gold[f'{name}:non_tree'])
    for key, got in zip(DET_KEYS, post_ref.get_detections_ref(gt, pred, 0.5, non_tree)):
        assert np.array_equal(got, gold[f'{name}:{key}']), key
    # the product's host half (counts -> matrices) on counts computed with numpy: bit-identical float64 matrices
    n_pred, n_gt = int(pred.max()) + 1, int(gt.max()) + 1
    mats = post.detection_matrices(_counts_numpy(pred, gt, n_pred, n_gt), non_tree)
    for key, got in zip(DET_KEYS[2:], mats):
        assert np.array_equal(got, gold[f'{name}:{key}']), key


def test_post_functions_need_the_gpu():
    if torch.cuda.is_available():
        pytest.skip('CPU-only check')
    from treelearn_b200._lib import TreeLearnCudaError
    with pytest.raises(TreeLearnCudaError):
        post.propagate_preds_hash_vox(np.zeros((2, 3)), np.zeros(2, np.int64), np.zeros((2, 3)))
    with pytest.raises(TreeLearnCudaError):
        post.get_detections(np.zeros(4, np.int64), np.zeros(4, np.int64), 0.5, -1)


# ---- GPU ------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize('name', KNN)
def test_propagate_preds_cuda_matches_reference_golden(gold, name):
    got = post.propagate_preds(gold[f'{name}:src'], gold[f'{name}:lab'], gold[f'{name}:tgt'], 5)
    assert got.dtype == np.int64 and np.array_equal(got, gold[f'{name}:out'])


@pytest.mark.gpu
def test_propagate_preds_cuda_too_few_sources_raises():
    with pytest.raises(ValueError):
        post.propagate_preds(np.zeros((3, 3)), np.zeros(3, np.int64), np.ones((10, 3)), 5)


@pytest.mark.gpu
@pytest.mark.parametrize('name', VOX)
def test_hash_vox_cuda_matches_reference_golden(gold, name):
    got, missing = post.propagate_preds_hash_vox(gold[f'{name}:cur'], gold[f'{name}:preds'], gold[f'{name}:ret'])
    assert got.dtype == np.int64 and np.array_equal(got, gold[f'{name}:out'])
    assert np.array_equal(missing, gold[f'{name}:missing'])


@pytest.mark.gpu
def test_hash_vox_cuda_edge_cases():
    # signed zeros are one key; a repeated key keeps the last prediction; nothing to look up / nothing to look up in
    cur = np.array([[-0.0, 0.004, 1.0], [2.0, 2.0, 2.0], [1.999, 2.001, 2.0], [7.0, 7.0, 7.0]])
    ret = np.array([[0.0, -0.0, 1.0], [2.0, 2.0, 2.0], [3.0, 3.0, 3.0]])
    got, missing = post.propagate_preds_hash_vox(cur, np.array([5, 6, 9, -1]), ret)
    assert got.tolist() == [5, 9, -1] and missing.tolist() == [False, False, True]
    got, missing = post.propagate_preds_hash_vox(cur, np.array([5, 6, 9, 4]), np.zeros((0, 3)))
    assert got.shape == (0,) and missing.shape == (0,)
    got, missing = post.propagate_preds_hash_vox(np.zeros((0, 3)), np.zeros(0, np.int64), ret)
    assert got.tolist() == [-1, -1, -1] and missing.all()
    # fp32 rows against fp64 rows compare as values: 0.5 and 0.25 are exact in both, 0.1 is not the same number
    got, _ = post.propagate_preds_hash_vox(np.array([[0.5, 0.25, 0.1]], np.float32), np.array([3]),
                                           np.array([[0.5, 0.25, 0.1], [0.5, 0.25, float(np.float32(0.1))]]))
    assert got.tolist() == [-1, 3]


@pytest.mark.gpu
def test_hash_vox_cuda_large_random_vs_oracle():
    rng = np.random.default_rng(7)
    ret = np.unique(rng.integers(0, 400, (300000, 3)), axis=0) / 100.0
    sel = rng.permutation(len(ret))[:200000]
    cur = ret[sel] + rng.uniform(-0.004, 0.004, (len(sel), 3))
    preds = rng.integers(0, 5000, len(sel))
    got, missing = post.propagate_preds_hash_vox(cur, preds, ret)
    want, want_missing = post_ref.propagate_preds_hash_vox_ref(cur, preds, ret)
    assert np.array_equal(got, want) and np.array_equal(missing, want_missing)
    assert 0 < missing.sum() < len(ret)


@pytest.mark.gpu
@pytest.mark.parametrize('name', DET)
def test_get_detections_cuda_matches_reference_golden(gold, name):
    gt, pred, non_tree = gold[f'{name}:gt'], gold[f'{name}:pred'], int(gold[f'{name}:non_tree'])
    for key, got in zip(DET_KEYS, post.get_detections(gt, pred, 0.5, non_tree)):
        assert np.array_equal(got, gold[f'{name}:{key}']), key


@pytest.mark.gpu
def test_cooccurrence_counts_cuda_large_vs_numpy():
    rng = np.random.default_rng(3)
    n, n_pred, n_gt = 3_000_000, 700, 650
    pred = np.sort(rng.integers(-2, n_pred + 2, n))          # long runs of equal labels, like a real plot
    gt = np.clip(pred + rng.integers(-1, 2, n), -1, n_gt + 1)
    got = post.cooccurrence_counts_cuda(torch.from_numpy(pred).cuda(), torch.from_numpy(gt).cuda(), n_pred, n_gt)
    assert np.array_equal(got.cpu().numpy(), _counts_numpy(pred, gt, n_pred, n_gt))
    assert int(got.sum()) == n
    empty = post.cooccurrence_counts_cuda(torch.zeros(0, dtype=torch.int64).cuda(), torch.zeros(0, dtype=torch.int64).cuda(), 3, 2)
    assert empty.shape == (4, 3) and int(empty.abs().sum()) == 0
