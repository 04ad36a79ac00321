"""ORACLE (test infrastructure): CPU restatement of the steps right after the per-tile path.

Follows /root/reference/tree_learn:
  propagate_preds_ref            util/pipeline.py:300-331  (fp32 coordinates, kNN, most frequent label via bincount.argmax
                                                            = smallest label among the most frequent)
  np_round2                      numpy.round(x, 2) as util/pipeline.py:443,458 apply it: multiply by 100, rint, divide by
                                 100 in the array's own dtype
  propagate_preds_hash_vox_ref   util/pipeline.py:455-465  (dict keyed by the coordinate triple; the reference keys by
                                 hash(tuple), which is the same map unless two different triples collide in 64 bits)
  get_detections_ref             util/eval.py:7-31, get_eval_components :230-238, get_segmentation_metrics :242-258
Pinned against the reference's OWN functions by tests/golden/make_golden_post.py -> tests/golden/post_small.npz.
numpy + scipy only.  Only tests/, smoke() and bench.py's cpu_baseline legs may import it.
"""
import numpy as np
import scipy.optimize
from scipy.spatial import cKDTree


def propagate_preds_ref(source_coords, source_preds, target_coords, n_neighbors):
    src = np.asarray(source_coords, dtype=np.float32).astype(np.float64)
    tgt = np.asarray(target_coords, dtype=np.float32).astype(np.float64)
    preds = np.asarray(source_preds).astype(np.int64)
    if n_neighbors > len(src):
        raise ValueError('Expected n_neighbors <= n_samples_fit')
    _, idx = cKDTree(src).query(tgt, k=n_neighbors)
    idx = idx.reshape(len(tgt), n_neighbors)
    out = np.empty(len(tgt), dtype=np.int64)
    for i, row in enumerate(preds[idx]):
        labels, counts = np.unique(row, return_counts=True)        # ascending labels: argmax -> smallest of the modes
        out[i] = labels[np.argmax(counts)]
    return out


def np_round2(a):
    a = np.asarray(a)
    if a.dtype == np.float32:
        return np.rint(a * np.float32(100.0)) / np.float32(100.0)
    a = a.astype(np.float64)
    return np.rint(a * 100.0) / 100.0


def _triples(a):
    return [tuple(float(v) for v in row) for row in np.asarray(a).reshape(-1, 3)]     # float(): -0.0 == 0.0 as dict keys


def propagate_preds_hash_vox_ref(coords, instance_preds, coords_to_return):
    table = {}
    for key, pred in zip(_triples(np_round2(coords)), np.asarray(instance_preds)):
        table[key] = int(pred)                                       # a repeated key keeps the last prediction
    out = np.array([table.get(key, -1) for key in _triples(coords_to_return)], dtype=np.int64).reshape(-1)
    return out, out == -1


def get_detections_ref(instance_labels, instance_preds, min_iou_match, non_tree_label):
    instance_labels, instance_preds = np.asarray(instance_labels), np.asarray(instance_preds)
    n_pred, n_gt = int(instance_preds.max()) + 1, int(instance_labels.max()) + 1
    iou, prec, rec = (np.zeros((n_pred, n_gt)) for _ in range(3))
    for p in range(n_pred):
        pm = instance_preds == p
        for g in np.unique(instance_labels[pm]):
            if g == non_tree_label:
                continue
            gm = instance_labels == g
            tp, fp, fn = int((pm & gm).sum()), int((pm & ~gm).sum()), int((~pm & gm).sum())
            iou[p, g] = tp / (tp + fp + fn)
            prec[p, g] = tp / (tp + fp)
            rec[p, g] = tp / (tp + fn)
    pre_p, pre_g = scipy.optimize.linear_sum_assignment(iou, maximize=True)
    keep = iou[pre_p, pre_g] > min_iou_match
    return pre_g[keep], pre_p[keep], iou, prec, rec
