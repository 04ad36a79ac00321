"""Golden vectors for the post-path rows: outputs of the REFERENCE's own `propagate_preds`,
`propagate_preds_hash_vox` (/root/reference/tree_learn/util/pipeline.py:300-331, 455-465) and `get_detections`
(/root/reference/tree_learn/util/eval.py:7-31) on seeded inputs.  Run in the build container:

    python tests/golden/make_golden_post.py        -> tests/golden/post_small.npz
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference  # noqa: E402

from oracle import post_ref  # noqa: E402


def knn_case(seed, n_src, n_tgt, n_labels):
    """A voxelised cloud with blocky labels (incl. -1 and 0) and denser target points around it."""
    rng = np.random.default_rng(seed)
    src = rng.uniform(0, 8, (n_src, 3))
    lab = (np.floor(src[:, 0] / 8 * n_labels).astype(np.int64) + 3 * np.floor(src[:, 1] / 4).astype(np.int64)) - 1
    flip = rng.random(n_src) < 0.15
    lab[flip] = rng.integers(-1, n_labels + 3, flip.sum())
    tgt = np.concatenate([src[rng.integers(0, n_src, n_tgt // 2)] + rng.normal(0, 0.05, (n_tgt // 2, 3)),
                          rng.uniform(-1, 9, (n_tgt - n_tgt // 2, 3))])
    return src, lab, tgt


def vox_case(seed, n, dtype):
    """coords_to_return on a 1 cm lattice; the current coords are a shuffled subset with sub-millimetre jitter, some
    rows repeated with conflicting predictions (last one wins), some predictions equal to -1, and signed zeros."""
    rng = np.random.default_rng(seed)
    lattice = np.unique(rng.integers(-300, 300, (n, 3)), axis=0)
    rng.shuffle(lattice)
    ret = (lattice / 100.0).astype(dtype)
    ret = post_ref.np_round2(ret)                     # what a saved, rounded voxel cloud holds in this dtype
    ret[0] = [0.0, -0.0, 0.5]
    keep = rng.permutation(len(ret))[: int(0.8 * len(ret))]
    cur = ret[keep].astype(dtype) + rng.uniform(-0.003, 0.003, (len(keep), 3)).astype(dtype)
    cur[0] = [-0.0, 0.001, 0.5]                       # rounds to (-0.0, 0.0, 0.5): the partner of ret[0] if it is kept
    preds = rng.integers(-1, 40, len(cur)).astype(np.int64)
    dup = rng.integers(0, len(cur), len(cur) // 10)
    cur = np.concatenate([cur, cur[dup]])
    preds = np.concatenate([preds, rng.integers(0, 40, len(dup))])
    order = rng.permutation(len(cur))
    extra = (rng.uniform(5, 6, (7, 3))).astype(dtype)                     # current points without any partner
    return np.concatenate([cur[order], extra]), np.concatenate([preds[order], np.arange(7)]), ret


def det_case(seed, n, n_gt, n_pred, non_tree):
    """Ground-truth trees as x-slabs, predictions as shifted slabs with label noise; -1 = non-tree on both sides."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(0, 1, n)
    gt = np.floor(x * n_gt).astype(np.int64)
    pred = np.floor(np.clip(x + 0.3 / n_pred, 0, 0.999999) * n_pred).astype(np.int64)
    noise = rng.random(n) < 0.1
    pred[noise] = rng.integers(-1, n_pred, noise.sum())
    gt[rng.random(n) < 0.05] = non_tree
    pred[pred == n_pred - 2] = n_pred - 3            # an empty prediction id in the middle of the range
    return gt, pred


def main():
    warnings.simplefilter('ignore')
    _, ref_pipeline, _ = import_reference()
    import tree_learn.util.eval as ref_eval
    assert ref_eval.__file__.startswith('/root/reference')
    out = {}
    for name, args in {'knn_a': (1, 4000, 3000, 6), 'knn_b': (2, 600, 2500, 3)}.items():
        src, lab, tgt = knn_case(*args)
        got = ref_pipeline.propagate_preds(src, lab, tgt, 5)
        assert np.array_equal(got, post_ref.propagate_preds_ref(src, lab, tgt, 5)), name
        out[f'{name}:src'], out[f'{name}:lab'], out[f'{name}:tgt'], out[f'{name}:out'] = src, lab, tgt, got
    for name, args in {'vox_f64': (3, 5000, np.float64), 'vox_f32': (4, 5000, np.float32)}.items():
        cur, preds, ret = vox_case(*args)
        got, missing = ref_pipeline.propagate_preds_hash_vox(cur, preds, ret)
        ora, ora_missing = post_ref.propagate_preds_hash_vox_ref(cur, preds, ret)
        assert np.array_equal(got, ora) and np.array_equal(missing, ora_missing), name
        print(name, 'rows', len(cur), '->', len(ret), 'missing', int(missing.sum()))
        out[f'{name}:cur'], out[f'{name}:preds'], out[f'{name}:ret'] = cur, preds, ret
        out[f'{name}:out'], out[f'{name}:missing'] = np.asarray(got, dtype=np.int64), missing
    for name, args in {'det_a': (5, 20000, 12, 14, -1), 'det_b': (6, 5000, 5, 4, 0)}.items():
        gt, pred = det_case(*args)
        res = ref_eval.get_detections(gt, pred, 0.5, args[4])
        ora = post_ref.get_detections_ref(gt, pred, 0.5, args[4])
        for r, o in zip(res, ora):
            assert np.array_equal(np.asarray(r), np.asarray(o)), name
        print(name, 'matched', len(res[0]), 'of', res[2].shape)
        out[f'{name}:gt'], out[f'{name}:pred'], out[f'{name}:non_tree'] = gt, pred, args[4]
        for key, r in zip(['matched_gts', 'matched_preds', 'iou', 'prec', 'rec'], res):
            out[f'{name}:{key}'] = np.asarray(r)
    path = os.path.join(HERE, 'post_small.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, f'{os.path.getsize(path) / 1024:.0f} KiB')


if __name__ == '__main__':
    main()
