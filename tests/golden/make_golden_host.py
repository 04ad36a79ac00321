"""Golden vectors for the host side of the drop-in boundary: outputs of the REFERENCE's own
`TreeDataset` in TRAINING mode (seeded augmentation, /root/reference/tree_learn/dataset/dataset.py:35-226), of its config
parser on its own YAML files (util/parser.py:23-41) and of its evaluation helpers (util/eval.py:35-260).  Run in the
build container (the reference is not on the GPU box):

    python tests/golden/make_golden_host.py        -> tests/golden/host_small.npz
"""
import json
import logging
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference, REF  # noqa: E402

from treelearn_b200 import synth  # noqa: E402

AUG = {'point_jitter': True, 'jitter': True, 'flip': True, 'rot': True, 'scaled': True}


def write_tiles(tmp):
    """Two tile files in the format the reference's SampleGenerator saves (points f32, feat f32 [n,1], instance_label i32,
    center f64 [3]); labels contain -1 (unlabeled), 0 (non-tree) and trees, one of them with fewer than 12 points."""
    paths = []
    for i, seed in enumerate((21, 22)):
        f = synth.synth_forest(edge=10.0, height=8.0, n_trees=3, seed=seed, ground_density=30.0)
        inst = f['inst'].astype(np.int32)
        inst[::17] = -1
        tiny = np.where(inst == inst.max())[0]
        inst[tiny[8:]] = 0                       # the last tree keeps 8 points (< 12: min() branch of getOffset)
        p = os.path.join(tmp, f'tile_{i}.npz')
        np.savez(p, points=f['coords'].astype(np.float32), feat=f['feat'].astype(np.float32).reshape(-1, 1),
                 instance_label=inst, center=np.array([1.5 * i, -2.0, 0.0]))
        paths.append(p)
    return paths


def main():
    import_reference()
    import tree_learn.dataset.dataset as ref_dataset
    import tree_learn.util.eval as ref_eval
    import tree_learn.util.parser as ref_parser
    assert ref_dataset.__file__.startswith(REF)
    logger = logging.getLogger('golden')
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        paths = write_tiles(tmp)
        for k in ('points', 'feat', 'instance_label', 'center'):
            for i, p in enumerate(paths):
                out[f'tile{i}:{k}'] = np.load(p)[k]
        ds = ref_dataset.TreeDataset(tmp, 8, True, logger, AUG)
        ds.data_paths = paths
        np.random.seed(1234)
        batch = ds.collate_fn([ds[0], ds[1], ds[0]])
        for key, val in batch.items():
            out[f'train_batch:{key}'] = val.numpy() if hasattr(val, 'numpy') else np.asarray(val)
        out['train_batch:next_random'] = np.random.rand(3)          # the RNG must be left in the same state
    # ---- config parser on the reference's own files (paths in default_args are relative to the reference root) ----
    # munch is absent here (MagicMock): call the function with a dict-returning stand-in for Munch.fromDict
    ref_parser.Munch = type('M', (), {'fromDict': staticmethod(lambda d: d)})
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        for name in ('configs/pipeline/pipeline.yaml', 'configs/training/train.yaml', 'configs/evaluation/evaluate.yaml'):
            if os.path.exists(name):
                out['config:' + name] = np.array(json.dumps(ref_parser.get_config(name), sort_keys=True))
    finally:
        os.chdir(cwd)
    # ---- evaluation helpers ----
    rng = np.random.default_rng(5)
    n = 4000
    coords = rng.uniform(0, 20, size=(n, 3))
    gt = (coords[:, 0] // 4).astype(np.int64)                      # 5 slabs = instances 0..4 (0 = non-tree)
    pred = gt.copy()
    flip = rng.random(n) < 0.15
    pred[flip] = rng.integers(0, 7, size=flip.sum())               # noise + two extra predicted ids (5, 6)
    pred[(gt == 3)] = 2                                            # under-segmentation: trees 2 and 3 share prediction 2
    from treelearn_b200 import post                                # noqa: F401  (get_detections needs a GPU; use the reference's)
    mg, mp, iou, prec, rec = ref_eval.get_detections(gt, pred, 0.5, 0)
    out.update({'eval:coords': coords, 'eval:gt': gt, 'eval:pred': pred, 'eval:matched_gts': mg, 'eval:matched_preds': mp,
                'eval:iou': iou, 'eval:prec': prec, 'eval:rec': rec})
    fails = ref_eval.get_detection_failures(mg, mp, np.unique(gt[gt != 0]), np.unique(pred[pred != 0]), iou, prec, rec, 0.5, 0.4)
    for name, v in zip(('non_matched_gts', 'non_matched_preds', 'pred_gt', 'gt_pred', 'gt_other'), fails):
        out['eval:fail:' + name] = np.asarray(v, dtype=np.float64)
    ident = {i: i for i in range(10)}
    no, xy, z = ref_eval.evaluate_instance_segmentation(pred, gt, mg, mp, coords, ident, ident, [0, 0.5, 1.0, 2.0], [0, 0.3, 0.7, 1.5])
    for name, df in (('no', no), ('xy', xy), ('z', z)):
        out[f'eval:{name}:columns'] = np.array(json.dumps(list(df.columns)))
        out[f'eval:{name}:values'] = df.to_numpy(dtype=np.float64)
    path = os.path.join(HERE, 'host_small.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, f'{os.path.getsize(path) / 1024:.0f} KiB')


if __name__ == '__main__':
    main()
