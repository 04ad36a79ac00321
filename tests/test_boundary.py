"""CPU tests of the drop-in boundary: the literal import statements of the reference's tools resolve
against this repo's `tree_learn` package, and the host-side helpers behind those names reproduce the golden vectors
recorded from the reference's own code (tests/golden/make_golden_host.py)."""
import json
import logging
import math
import os
import re
import tempfile

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, 'tests', 'golden', 'host_small.npz'), allow_pickle=False)

# the import blocks of the reference's entry points, verbatim (file:line in /root/reference)
TOOL_IMPORTS = {
    'tools/pipeline/pipeline.py:7-13': '''
from tree_learn.dataset import TreeDataset
from tree_learn.model import TreeLearn
from tree_learn.util import (munch_to_dict, build_dataloader, get_root_logger, load_checkpoint, ensemble,
                             get_coords_within_shape, get_hull_buffer, get_hull, get_cluster_means,
                             propagate_preds, save_treewise, load_data, save_data, make_labels_consecutive,
                             get_config, generate_tiles, assign_remaining_points_nearest_neighbor,
                             get_pointwise_preds, get_instances, propagate_preds_hash_full, propagate_preds_hash_vox)
''',
    'tools/training/train.py:8-12': '''
from tree_learn.util import (checkpoint_save, init_train_logger, load_checkpoint,
                            is_multiple, get_args_and_cfg, build_cosine_scheduler, build_optimizer,
                            point_wise_loss, get_eval_components, build_dataloader)
from tree_learn.model import TreeLearn
from tree_learn.dataset import TreeDataset
''',
    'tools/evaluation/evaluate.py:7-9': '''
from tree_learn.util import (get_root_logger, make_labels_consecutive, get_config,
                             get_detections, get_detection_failures, save_data,
                             evaluate_instance_segmentation, propagate_preds, load_data)
''',
    'tools/data_gen/gen_train_data.py:7': '''
from tree_learn.util import SampleGenerator, get_root_logger, get_config, voxelize, compute_features, load_data
''',
    'tools/data_gen/gen_val_data.py:4': '''
from tree_learn.util import get_root_logger, get_config, generate_tiles
''',
}


@pytest.mark.parametrize('where', sorted(TOOL_IMPORTS))
def test_reference_tool_imports_resolve(where):
    ns = {}
    exec(TOOL_IMPORTS[where], ns)          # raises ImportError / AttributeError if a name is missing
    import tree_learn
    assert os.path.dirname(tree_learn.__file__) == os.path.join(ROOT, 'tree_learn')
    names = set(re.findall(r'\b([A-Za-z_][A-Za-z0-9_]*)\b', TOOL_IMPORTS[where].replace('from tree_learn', '').replace('import', '')))
    for n in names - {'dataset', 'model', 'util', 'tree_learn'}:
        assert callable(ns[n]), n


def test_tree_dataset_training_mode_matches_reference():
    """Same files, same np.random seed -> the batch the reference's TreeDataset + collate_fn produce, bit for bit
    (augmentation draw order, offsets, masks, dtypes), and the RNG left in the same state."""
    from tree_learn.dataset import TreeDataset
    aug = {'point_jitter': True, 'jitter': True, 'flip': True, 'rot': True, 'scaled': True}
    with tempfile.TemporaryDirectory() as tmp:
        paths = []
        for i in range(2):
            p = os.path.join(tmp, f'tile_{i}.npz')
            np.savez(p, **{k: GOLD[f'tile{i}:{k}'] for k in ('points', 'feat', 'instance_label', 'center')})
            paths.append(p)
        ds = TreeDataset(tmp, 8, True, logging.getLogger('t'), aug)
        assert len(ds) == 2
        ds.data_paths = paths
        np.random.seed(1234)
        batch = ds.collate_fn([ds[0], ds[1], ds[0]])
        nxt = np.random.rand(3)
    assert np.array_equal(nxt, GOLD['train_batch:next_random'])
    keys = [k.split(':', 1)[1] for k in GOLD.files if k.startswith('train_batch:') and not k.endswith('next_random')]
    assert set(keys) == set(batch), set(keys) ^ set(batch)
    for k in keys:
        want = GOLD['train_batch:' + k]
        got = batch[k].numpy() if torch.is_tensor(batch[k]) else np.asarray(batch[k])
        assert got.dtype == want.dtype and got.shape == want.shape, (k, got.dtype, want.dtype)
        assert np.array_equal(got, want), k


def test_get_config_matches_reference_on_its_own_yaml_files():
    from tree_learn.util import get_config, munch_to_dict
    cfgs = [k for k in GOLD.files if k.startswith('config:')]
    assert cfgs
    with tempfile.TemporaryDirectory() as tmp:
        # re-create the reference's config tree from the golden (include paths are relative to the working directory)
        want = {k[7:]: json.loads(str(GOLD[k])) for k in cfgs}
        mod = os.path.join(tmp, 'configs', '_modular')
        os.makedirs(mod)
        import yaml
        full = want['configs/pipeline/pipeline.yaml']
        # split the merged config back into one include + a main file overriding one nested key
        include = {'model': dict(full['model'], spatial_shape=None), 'grouping': full['grouping']}
        with open(os.path.join(mod, 'inc.yaml'), 'w') as f:
            yaml.safe_dump(include, f)
        main = {k: v for k, v in full.items() if k not in ('grouping',)}
        main['model'] = {'spatial_shape': full['model']['spatial_shape']}
        main['default_args'] = ['configs/_modular/inc.yaml']
        with open(os.path.join(tmp, 'main.yaml'), 'w') as f:
            yaml.safe_dump(main, f)
        cwd = os.getcwd()
        os.chdir(tmp)
        try:
            cfg = get_config('main.yaml')
        finally:
            os.chdir(cwd)
    assert munch_to_dict(cfg) == full
    assert cfg.model.spatial_shape == full['model']['spatial_shape'] and cfg.grouping.tau_min == full['grouping']['tau_min']


def test_eval_helpers_match_reference():
    from tree_learn.util import evaluate_instance_segmentation, get_detection_failures, get_eval_components
    from treelearn_b200.host_util import get_segmentation_metrics
    gt, pred, coords = GOLD['eval:gt'], GOLD['eval:pred'], GOLD['eval:coords']
    mg, mp = GOLD['eval:matched_gts'], GOLD['eval:matched_preds']
    iou, prec, rec = GOLD['eval:iou'], GOLD['eval:prec'], GOLD['eval:rec']
    # matrices entry by entry from the mask helpers
    for p in range(iou.shape[0]):
        for g in range(1, iou.shape[1]):
            tp, fp, tn, fn = get_eval_components(pred == p, gt == g)
            if tp == 0:
                continue
            pr, rc, io = get_segmentation_metrics(tp, fp, fn)
            assert (pr, rc, io) == (prec[p, g], rec[p, g], iou[p, g])
    fails = get_detection_failures(mg, mp, np.unique(gt[gt != 0]), np.unique(pred[pred != 0]), iou, prec, rec, 0.5, 0.4)
    for name, v in zip(('non_matched_gts', 'non_matched_preds', 'pred_gt', 'gt_pred', 'gt_other'), fails):
        assert np.array_equal(np.asarray(v, dtype=np.float64), GOLD['eval:fail:' + name], equal_nan=True), name
    ident = {i: i for i in range(10)}
    tables = evaluate_instance_segmentation(pred, gt, mg, mp, coords, ident, ident, [0, 0.5, 1.0, 2.0], [0, 0.3, 0.7, 1.5])
    for name, df in zip(('no', 'xy', 'z'), tables):
        assert list(df.columns) == json.loads(str(GOLD[f'eval:{name}:columns'])), name
        assert np.array_equal(df.to_numpy(dtype=np.float64), GOLD[f'eval:{name}:values'], equal_nan=True), name


def test_cosine_schedule_closed_form():
    """timm's CosineLRScheduler as the reference configures it (configs/training/train.yaml: t_initial 1300, lr_min 1e-4,
    cycle_decay 1, warmup_lr_init 2e-5, warmup_t 50, cycle_limit 1, t_in_epochs True)."""
    from types import SimpleNamespace
    from tree_learn.util import build_cosine_scheduler
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.AdamW([p], lr=3e-3)
    cfg = SimpleNamespace(t_initial=1300, lr_min=1e-4, cycle_decay=1, warmup_lr_init=2e-5, warmup_t=50, cycle_limit=1, t_in_epochs=True)
    sch = build_cosine_scheduler(cfg, opt)
    assert opt.param_groups[0]['lr'] == 2e-5                      # warm-up start value set at construction
    for epoch, want in [(0, 2e-5), (25, 2e-5 + 25 * (3e-3 - 2e-5) / 50), (50, 1e-4 + 0.5 * (3e-3 - 1e-4) * (1 + math.cos(math.pi * 50 / 1300))),
                        (650, 1e-4 + 0.5 * (3e-3 - 1e-4)), (1299, 1e-4 + 0.5 * (3e-3 - 1e-4) * (1 + math.cos(math.pi * 1299 / 1300))), (1300, 1e-4), (2000, 1e-4)]:
        sch.step(epoch)
        assert opt.param_groups[0]['lr'] == pytest.approx(want, rel=1e-12), epoch


def test_point_cloud_files_round_trip_and_cluster_means():
    import pandas as pd
    from tree_learn.util import get_cluster_means, load_data, save_data
    rng = np.random.default_rng(0)
    data = np.hstack([rng.normal(size=(50, 3)), rng.integers(0, 4, size=(50, 1)).astype(np.float64)])
    with tempfile.TemporaryDirectory() as tmp:
        for fmt in ('npy', 'npz'):
            save_data(data, fmt, 'cloud', tmp)
            assert np.array_equal(load_data(os.path.join(tmp, f'cloud.{fmt}')), data)
        save_data(data, 'txt', 'cloud', tmp)
        assert np.allclose(load_data(os.path.join(tmp, 'cloud.txt')), data[1:])       # read_csv takes the first row as header
        np.save(os.path.join(tmp, 'xyz.npy'), data[:, :3])
        assert np.array_equal(load_data(os.path.join(tmp, 'xyz.npy'))[:, 3], -np.ones(50))
        with pytest.raises(ImportError):
            save_data(data, 'laz', 'cloud', tmp)          # laspy is not installed: raises on CALL, not on import
    coords, labels = data[:, :3].astype(np.float32), data[:, 3].astype(np.int64)
    df = pd.DataFrame(coords, columns=['x', 'y', 'z'])
    df['label'] = labels
    assert np.allclose(get_cluster_means(coords, labels), df.groupby('label').mean().values, rtol=1e-6)


def test_plot_outline_known_answers():
    from tree_learn.util import get_coords_within_shape, get_hull, get_hull_buffer
    g = np.arange(0, 10.01, 0.2)
    xx, yy = np.meshgrid(g, g)
    square = np.stack([xx.ravel(), yy.ravel()], 1) + np.array([1000.0, -500.0])
    # an L-shaped plot: the alpha shape must follow the notch, the convex hull (alpha = 0) must not
    lshape = square[~((square[:, 0] > 1005.1) & (square[:, 1] > -494.9))]
    probe = np.array([[1002.0, -498.0, 0.0], [1006.5, -493.5, 0.0], [1008.0, -498.0, 0.0], [1011.0, -498.0, 0.0]])
    assert get_coords_within_shape(probe, get_hull(lshape, 0.6)).tolist() == [True, False, True, False]
    assert get_coords_within_shape(probe, get_hull(lshape, 0)).tolist() == [True, True, True, False]
    edge = np.array([[1000.1, -495.0, 0.0], [1003.0, -495.0, 0.0], [1005.05, -493.0, 0.0], [1009.9, -499.95, 0.0]])
    assert get_coords_within_shape(edge, get_hull_buffer(lshape, 0.6, 0.3)).tolist() == [True, False, True, True]
