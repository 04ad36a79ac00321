"""The steps right after the per-tile path, same names / arguments / return values as the
reference, numpy in and numpy out, computed by the CUDA library:

  propagate_preds            <- tree_learn/util/pipeline.py:300-331   (kNN majority vote, reuses tl_knn_vote)
  propagate_preds_hash_vox   <- tree_learn/util/pipeline.py:455-465   (exact-coordinate join, tl_hash_join_last)
  get_detections             <- tree_learn/util/eval.py:7-31          (tl_cooccurrence_counts + scipy Hungarian matching)

`propagate_preds_hash_full` (util/pipeline.py:441-452) is keyed by a dictionary of *Python hash values* that the
reference's tile generation builds (`get_hash_mapping`, fed by open3d's voxel_down_sample_and_trace); it belongs with
that step (§8f row 2) and is not provided here.
"""
import numpy as np
import scipy.optimize
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr
from .pipeline import _dev, _ws, knn_vote_cuda


def propagate_preds(source_coords, source_preds, target_coords, n_neighbors, n_jobs=1):
    """Label of every target point = most frequent label among its `n_neighbors` nearest source points (fp32
    coordinates as in the reference; ties -> smallest label, which is what bincount().argmax() returns)."""
    dev = _dev()
    source_coords = np.ascontiguousarray(source_coords, dtype=np.float32)
    target_coords = np.ascontiguousarray(target_coords, dtype=np.float32)
    source_preds = np.ascontiguousarray(source_preds).astype(np.int64)
    if n_neighbors > len(source_coords):       # sklearn's kneighbors raises the same way
        raise ValueError(f'Expected n_neighbors <= n_samples_fit, but n_neighbors = {n_neighbors}, '
                         f'n_samples_fit = {len(source_coords)}, n_samples = {len(target_coords)}')
    out = knn_vote_cuda(torch.from_numpy(source_coords).to(dev), torch.from_numpy(source_preds).to(dev),
                        torch.from_numpy(target_coords).to(dev), n_neighbors)
    return out.cpu().numpy()


def _rows(a):
    """[n,3] coordinates as a contiguous fp32 or fp64 array (other dtypes are promoted to fp64, like Python floats)."""
    a = np.asarray(a)
    if a.dtype != np.float32:
        a = a.astype(np.float64, copy=False)
    return np.ascontiguousarray(a.reshape(-1, 3))


def hash_join_last_cuda(build_xyz, build_round2, build_vals, probe_xyz, probe_round2, missing=-1):
    """Device-resident join: build_xyz [B,3] / probe_xyz [P,3] f32 or f64 CUDA tensors, build_vals [B] i64 ->
    [P] i64 (value of the last build row with the probe row's coordinates, else `missing`)."""
    lib = _lib.load()
    nb, npb = int(build_xyz.shape[0]), int(probe_xyz.shape[0])
    out = torch.empty(npb, dtype=torch.int64, device=probe_xyz.device)
    if npb == 0:
        return out
    for t in (build_xyz, probe_xyz):
        assert t.dtype in (torch.float32, torch.float64) and t.is_contiguous()
    build_vals = build_vals.contiguous().long()
    wsb = lib.tl_hash_join_workspace_bytes(nb)
    ws = _ws(wsb, probe_xyz.device)
    check(lib.tl_hash_join_last(ptr(build_xyz), int(build_xyz.dtype == torch.float64), int(build_round2),
                                ptr(build_vals), nb, ptr(probe_xyz), int(probe_xyz.dtype == torch.float64),
                                int(probe_round2), npb, int(missing), ptr(out), ptr(ws), wsb, stream_ptr()))
    return out


def propagate_preds_hash_vox(coords, instance_preds, coords_to_return):
    """Predictions of the (2-decimal-rounded) `coords` looked up at `coords_to_return` by exact coordinates; -1 where
    a point has no partner.  Returns (preds_to_return, not_yet_propagated) like the reference -- including its
    convention that a propagated prediction equal to -1 also counts as "not yet propagated"."""
    dev = _dev()
    cur = torch.from_numpy(_rows(coords)).to(dev)
    ret = torch.from_numpy(_rows(coords_to_return)).to(dev)
    vals = torch.from_numpy(np.ascontiguousarray(instance_preds).astype(np.int64)).to(dev)
    assert vals.shape[0] == cur.shape[0]
    preds = hash_join_last_cuda(cur, True, vals, ret, False, missing=-1).cpu().numpy()
    return preds, preds == -1


def cooccurrence_counts_cuda(instance_preds, instance_labels, n_pred, n_gt):
    """[n_pred+1, n_gt+1] int64 point counts (CUDA tensors in and out); see tl_cooccurrence_counts."""
    lib = _lib.load()
    instance_preds, instance_labels = instance_preds.contiguous().long(), instance_labels.contiguous().long()
    assert instance_preds.shape == instance_labels.shape
    counts = torch.empty((n_pred + 1, n_gt + 1), dtype=torch.int64, device=instance_preds.device)
    check(lib.tl_cooccurrence_counts(ptr(instance_preds), ptr(instance_labels), int(instance_preds.numel()), int(n_pred),
                                     int(n_gt), ptr(counts), stream_ptr()))
    return counts


def detection_matrices(counts, non_tree_label):
    """iou / precision / recall matrices [n_pred, n_gt] (float64) from the co-occurrence counts, entry by entry what
    get_eval_components + get_segmentation_metrics give: filled only where a prediction and a label share points
    (tree_learn/util/eval.py:13-24), the `non_tree_label` column left at zero."""
    counts = np.asarray(counts, dtype=np.int64)
    tp = counts[:-1, :-1]
    n_p = counts.sum(axis=1)[:-1, None]          # |pred == p|  = tp + fp
    n_g = counts.sum(axis=0)[None, :-1]          # |label == g| = tp + fn
    hit = tp > 0
    if 0 <= non_tree_label < tp.shape[1]:
        hit = hit.copy()
        hit[:, non_tree_label] = False
    with np.errstate(divide='ignore', invalid='ignore'):
        iou = np.where(hit, tp / (n_p + n_g - tp), 0.0)
        prec = np.where(hit, tp / np.broadcast_to(n_p, tp.shape), 0.0)
        rec = np.where(hit, tp / np.broadcast_to(n_g, tp.shape), 0.0)
    return iou, prec, rec


def get_detections(instance_labels, instance_preds, min_iou_match, non_tree_label):
    """Hungarian matching of predicted and ground-truth instances on the IoU matrix.  Returns (matched_gts,
    matched_preds, iou_matrix, precision_matrix, recall_matrix) exactly as the reference does."""
    dev = _dev()
    instance_labels = np.asarray(instance_labels)
    instance_preds = np.asarray(instance_preds)
    n_pred, n_gt = int(np.max(instance_preds)) + 1, int(np.max(instance_labels)) + 1
    stray = (instance_labels < 0) & (instance_labels != non_tree_label)
    if stray.any():          # the reference would index its matrices from the end with such a label
        raise ValueError('get_detections: negative instance labels other than non_tree_label are not supported')
    counts = cooccurrence_counts_cuda(torch.from_numpy(instance_preds.astype(np.int64)).to(dev),
                                      torch.from_numpy(instance_labels.astype(np.int64)).to(dev), n_pred, n_gt)
    iou, prec, rec = detection_matrices(counts.cpu().numpy(), non_tree_label)
    pre_p, pre_g = scipy.optimize.linear_sum_assignment(iou, maximize=True)
    keep = iou[pre_p, pre_g] > min_iou_match
    return pre_g[keep], pre_p[keep], iou, prec, rec
