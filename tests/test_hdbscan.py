"""HDBSCAN row (the reference's DEFAULT clusterer): the oracle restatement, the library's host-side
dendrogram / condensed-tree pass and the CUDA core-distance + Prim-MST kernels, all against golden vectors recorded from
the reference's own `group_hdbscan` / `get_instances(use_hdbscan=True)` (tests/golden/make_golden_hdbscan.py)."""
import ctypes as C
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import cluster_ref
from treelearn_b200 import _lib

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
CASES = ['a', 'b', 'c', 'd']


@pytest.fixture(scope='module')
def gold():
    return np.load(os.path.join(GOLD, 'hdbscan_small.npz'))


def _case(g, name):
    return g[f'{name}:points'], int(g[f'{name}:mcs']), g[f'{name}:sklearn_labels'], g[f'{name}:group_hdbscan']


# ---- CPU ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('name', ['c', 'd'])
def test_oracle_hdbscan_matches_sklearn_golden(gold, name):
    pts, mcs, raw, grouped = _case(gold, name)
    assert np.array_equal(cluster_ref.hdbscan_ref(pts, mcs), raw)
    assert np.array_equal(cluster_ref.group_hdbscan_ref(pts, mcs, -1, 1), grouped)


@pytest.mark.parametrize('name', CASES)
def test_host_tree_labels_match_sklearn_golden(gold, name):
    """tl_hdbscan_tree_labels is host code of the C-ABI library: fed with the oracle's MST it must reproduce sklearn."""
    pts, mcs, raw, _ = _case(gold, name)
    lib = C.CDLL(_lib.LIB_PATH)
    lib.tl_hdbscan_tree_labels.argtypes = [C.c_void_p] * 3 + [C.c_int64, C.c_int64, C.c_void_p]
    core = cluster_ref.core_distances_ref(pts, mcs)
    src, dst, w = cluster_ref.prim_mst_ref(pts, core)
    order = np.argsort(w)
    s, d, ww = (np.ascontiguousarray(a[order]) for a in (src, dst, w))
    labels = np.empty(len(pts), dtype=np.int64)
    assert lib.tl_hdbscan_tree_labels(s.ctypes.data, d.ctypes.data, ww.ctypes.data, len(pts), mcs, labels.ctypes.data) == 0
    assert np.array_equal(labels, raw)


# ---- GPU ------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize('name', CASES)
def test_core_distance_and_mst_bit_exact_vs_oracle(gold, name):
    from treelearn_b200._lib import check, ptr, stream_ptr
    pts, mcs, _, _ = _case(gold, name)
    lib = _lib.load()
    n = len(pts)
    p = torch.from_numpy(pts).cuda()
    core = torch.empty(n, dtype=torch.float64, device='cuda')
    src = torch.empty(n - 1, dtype=torch.int32, device='cuda')
    dst = torch.empty(n - 1, dtype=torch.int32, device='cuda')
    w = torch.empty(n - 1, dtype=torch.float64, device='cuda')
    wsb = lib.tl_hdbscan_workspace_bytes(n)
    ws = torch.empty(wsb, dtype=torch.uint8, device='cuda')
    check(lib.tl_core_distance(ptr(p), n, mcs, ptr(core), ptr(ws), wsb, stream_ptr()))
    check(lib.tl_mst_prim(ptr(p), ptr(core), n, ptr(src), ptr(dst), ptr(w), ptr(ws), wsb, stream_ptr()))
    core_ref = cluster_ref.core_distances_ref(pts, mcs)
    assert np.array_equal(core.cpu().numpy(), core_ref)                 # fp64, no FMA: bit exact
    s_ref, d_ref, w_ref = cluster_ref.prim_mst_ref(pts, core_ref)
    assert np.array_equal(w.cpu().numpy(), w_ref)
    assert np.array_equal(dst.cpu().numpy().astype(np.int64), d_ref)     # same insertion order, same tie breaks
    assert np.array_equal(src.cpu().numpy().astype(np.int64), s_ref)


@pytest.mark.gpu
@pytest.mark.parametrize('name', CASES)
def test_group_hdbscan_matches_reference_golden(gold, name):
    from treelearn_b200 import pipeline
    pts, mcs, raw, grouped = _case(gold, name)
    assert np.array_equal(pipeline.hdbscan_cuda(torch.from_numpy(pts).cuda(), mcs), raw)
    out = pipeline.group_hdbscan(pts, mcs, -1, 1)
    assert out.dtype == np.int64 and np.array_equal(out, grouped)


@pytest.mark.gpu
def test_get_instances_with_default_clusterer_matches_reference_golden(gold):
    from treelearn_b200 import pipeline
    g = np.load(os.path.join(GOLD, 'cluster_small.npz'))
    cfg = SimpleNamespace(tree_conf_thresh=0.5, tau_vert=0.6, tau_off=4, tau_group=0.15, tau_min=50, use_hdbscan=True)
    inst = pipeline.get_instances(g['ens_out:coords'], g['ens_out:offset_predictions'], g['ens_out:semantic_scores'], cfg,
                                  g['ens_out:input_feats'][:, -1], 0, 0, -1, 1)
    assert np.array_equal(inst, gold['instances_hdbscan'])


@pytest.mark.gpu
def test_hdbscan_sparse_points_take_the_exact_scan_and_errors_are_sklearns():
    """Isolated points exceed the grid search radius (exact fallback scan); sample-count errors read like sklearn's."""
    from treelearn_b200 import pipeline
    rng = np.random.default_rng(0)
    pts = np.concatenate([rng.normal(0, 0.1, (200, 2)), rng.uniform(-500, 500, (60, 2))]).astype(np.float32)
    assert np.array_equal(pipeline.hdbscan_cuda(torch.from_numpy(pts).cuda(), 20), cluster_ref.hdbscan_ref(pts, 20))
    with pytest.raises(ValueError, match='must be at most the number of samples'):
        pipeline.hdbscan_cuda(torch.zeros((10, 2), device='cuda'), 50)
