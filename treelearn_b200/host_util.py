"""Host-side helpers the reference's tools import from `tree_learn.util` beside the hot path -- configuration, logging,
LR schedule, evaluation metrics, point-cloud I/O, plot hulls, file-based tile generation -- so that
`tools/pipeline/pipeline.py`, `tools/training/train.py`, `tools/evaluation/evaluate.py` and `tools/data_gen/gen_val_data.py`
import and run against this package unchanged.  None of this is on the GPU hot path; it is plain
numpy / torch written for this package, each function naming the reference function whose contract it keeps.

Optional third-party packages of the reference (`munch`, `tensorboardX`, `timm`, `laspy`) are used when installed and
replaced by small equivalents otherwise; what cannot work without them (LAS files) raises on CALL, never on import.
"""
import argparse
import json
import logging
import math
import os
import os.path as osp
import pickle
import random
import shutil
import time

import numpy as np
import torch
import yaml

INSTANCE_LABEL_IGNORE_IN_RAW_DATA = -1
NON_TREE_CLASS_IN_RAW_DATA = 0

# ---- configuration (tree_learn/util/parser.py) ----------------------------------------------------------------------
try:
    from munch import Munch
except ImportError:
    class Munch(dict):
        """Attribute-access dictionary (the subset of `munch.Munch` the tools use)."""

        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)

        def __setattr__(self, k, v):
            self[k] = v

        def __delattr__(self, k):
            try:
                del self[k]
            except KeyError:
                raise AttributeError(k)

        @classmethod
        def fromDict(cls, d):
            if isinstance(d, dict):
                return cls((k, cls.fromDict(v)) for k, v in d.items())
            if isinstance(d, (list, tuple)):
                return type(d)(cls.fromDict(v) for v in d)
            return d

        def toDict(self):
            return munch_to_dict(self)

        def copy(self):
            return type(self)(self)


def get_args(args):
    """parser.py:6-16."""
    parser = argparse.ArgumentParser('tree_learn')
    parser.add_argument('--config', type=str, help='path to config file')
    parser.add_argument('--resume', type=str, help='path to resume from')
    parser.add_argument('--work_dir', type=str, help='working directory')
    parser.add_argument('--dist', action='store_true', help='distributed training')
    return parser.parse_args() if args is None else parser.parse_args(args)


def load_yaml_file(filepath):
    with open(filepath, 'r') as file:
        return yaml.safe_load(file)


def modify_default_cfg(default_config, main_cfg):
    """Deep-merge the main file's entries over an included default file (parser.py:55-60)."""
    for key, value in main_cfg.items():
        if isinstance(value, dict) and isinstance(default_config.get(key), dict):
            modify_default_cfg(default_config[key], value)
        else:
            default_config[key] = value


def get_config(config_path):
    """YAML with `default_args: [paths]` includes (parser.py:23-41): keys present in the main file are merged over the
    include, then the include is written into the main config."""
    main_cfg = load_yaml_file(config_path)
    for path in main_cfg.pop('default_args', None) or []:
        default_config = load_yaml_file(path)
        for key in main_cfg:
            if key in default_config:
                modify_default_cfg(default_config[key], main_cfg[key])
        main_cfg.update(default_config)
    return Munch.fromDict(main_cfg)


def get_args_and_cfg(args=None):
    args = get_args(args)
    cfg = get_config(args.config)
    print(args)
    cfg.work_dir = osp.join('./work_dirs', args.work_dir if args.work_dir is not None else osp.splitext(osp.basename(args.config))[0])
    return args, cfg


def munch_to_dict(obj):
    if isinstance(obj, Munch):
        return {key: munch_to_dict(value) for key, value in obj.items()}
    if isinstance(obj, list):
        return [munch_to_dict(item) for item in obj]
    if isinstance(obj, tuple):
        return tuple(munch_to_dict(item) for item in obj)
    return obj


# ---- logging (tree_learn/util/logger.py) ----------------------------------------------------------------------------
def get_root_logger(log_file=None, log_level=logging.INFO):
    logger = logging.getLogger('TreeLearn')
    if logger.hasHandlers():
        return logger
    logging.basicConfig(format='%(asctime)s - %(levelname)s - %(message)s', level=log_level)
    if log_file is not None:
        handler = logging.FileHandler(log_file, 'w')
        handler.setFormatter(logging.Formatter('%(asctime)s - %(levelname)s - %(message)s'))
        handler.setLevel(log_level)
        logger.addHandler(handler)
    return logger


def _summary_writer_base():
    try:
        from tensorboardX import SummaryWriter as base
        return base
    except ImportError:
        pass
    try:
        from torch.utils.tensorboard import SummaryWriter as base
        return base
    except Exception:
        pass

    class JsonlWriter:
        """Scalars as JSON lines in `<logdir>/scalars.jsonl` when no tensorboard writer is installed."""

        def __init__(self, logdir, *args, **kwargs):
            os.makedirs(logdir, exist_ok=True)
            self._f = open(os.path.join(logdir, 'scalars.jsonl'), 'a')

        def add_scalar(self, tag, value, global_step=None, *args, **kwargs):
            self._f.write(json.dumps({'tag': tag, 'value': float(value), 'step': global_step}) + '\n')

        def flush(self, *args, **kwargs):
            self._f.flush()

        def close(self):
            self._f.close()
    return JsonlWriter


class SummaryWriter:
    """`tree_learn.util.SummaryWriter`: add_scalar / flush on whichever writer backend exists."""

    def __init__(self, *args, **kwargs):
        self._w = _summary_writer_base()(*args, **kwargs)

    def add_scalar(self, *args, **kwargs):
        return self._w.add_scalar(*args, **kwargs)

    def flush(self, *args, **kwargs):
        return self._w.flush(*args, **kwargs)

    def __getattr__(self, name):
        return getattr(self._w, name)


def init_train_logger(cfg, args):
    """logger.py:35-45: work_dir, timestamped log file, config copy, scalar writer."""
    os.makedirs(os.path.abspath(cfg.work_dir), exist_ok=True)
    log_file = os.path.join(cfg.work_dir, f'{time.strftime("%Y%m%d_%H%M%S", time.localtime())}.log')
    logger = get_root_logger(log_file=log_file)
    logger.info(f'Config:\n{cfg}')
    logger.info(f'Mix precision training: {cfg.fp16}')
    shutil.copy(args.config, os.path.join(cfg.work_dir, os.path.basename(args.config)))
    return logger, SummaryWriter(cfg.work_dir)


# ---- LR schedule (tree_learn/util/train.py:113-122 -> timm.scheduler.CosineLRScheduler) -------------------------------
class CosineSchedule:
    """The subset of timm's `CosineLRScheduler` the reference configures (no noise, cycle_mul = 1, k_decay = 1):
    linear warm-up from `warmup_lr_init` over `warmup_t` steps, then lr_min + (lr_max * cycle_decay^i - lr_min) / 2 *
    (1 + cos(pi * t_cur / t_initial)) in cycle i, lr_min after `cycle_limit` cycles.  `step(epoch)` / `step_update(n)`
    act when the unit matches `t_in_epochs`, like timm."""

    def __init__(self, optimizer, t_initial, lr_min=0.0, cycle_decay=1.0, warmup_lr_init=0.0, warmup_t=0, cycle_limit=1,
                 t_in_epochs=True, warmup_prefix=False):
        self.optimizer, self.t_initial, self.lr_min, self.cycle_decay = optimizer, t_initial, lr_min, cycle_decay
        self.warmup_lr_init, self.warmup_t, self.cycle_limit = warmup_lr_init, warmup_t, cycle_limit
        self.t_in_epochs, self.warmup_prefix = t_in_epochs, warmup_prefix
        for group in optimizer.param_groups:
            group.setdefault('initial_lr', group['lr'])
        self.base_values = [group['initial_lr'] for group in optimizer.param_groups]
        if warmup_t:
            self._set([warmup_lr_init for _ in self.base_values])

    def _set(self, values):
        for group, v in zip(self.optimizer.param_groups, values):
            group['lr'] = v

    def _get_lr(self, t):
        if t < self.warmup_t:
            return [self.warmup_lr_init + t * (b - self.warmup_lr_init) / self.warmup_t for b in self.base_values]
        if self.warmup_prefix:
            t = t - self.warmup_t
        i = t // self.t_initial
        t_cur = t - self.t_initial * i
        if i >= self.cycle_limit:
            return [self.lr_min for _ in self.base_values]
        gamma = self.cycle_decay ** i
        return [self.lr_min + 0.5 * (b * gamma - self.lr_min) * (1 + math.cos(math.pi * t_cur / self.t_initial))
                for b in self.base_values]

    def step(self, epoch, metric=None):
        if self.t_in_epochs:
            self._set(self._get_lr(epoch))

    def step_update(self, num_updates, metric=None):
        if not self.t_in_epochs:
            self._set(self._get_lr(num_updates))

    def state_dict(self):
        return {k: v for k, v in self.__dict__.items() if k != 'optimizer'}

    def load_state_dict(self, state):
        self.__dict__.update(state)


def build_cosine_scheduler(cfg, optimizer):
    kwargs = dict(t_initial=cfg.t_initial, lr_min=cfg.lr_min, cycle_decay=cfg.cycle_decay, warmup_lr_init=cfg.warmup_lr_init,
                  warmup_t=cfg.warmup_t, cycle_limit=cfg.cycle_limit, t_in_epochs=cfg.t_in_epochs)
    try:
        from timm.scheduler.cosine_lr import CosineLRScheduler
        return CosineLRScheduler(optimizer, **kwargs)
    except ImportError:
        return CosineSchedule(optimizer, **kwargs)


# ---- evaluation (tree_learn/util/eval.py) ------------------------------------------------------------------------------
def get_eval_components(preds_mask, labels_mask):
    """tp, fp, tn, fn of two boolean masks (eval.py:230-238)."""
    assert len(preds_mask) == len(labels_mask)
    not_p, not_l = np.logical_not(preds_mask), np.logical_not(labels_mask)
    return (preds_mask & labels_mask).sum(), (preds_mask & not_l).sum(), (not_p & not_l).sum(), (not_p & labels_mask).sum()


def get_segmentation_metrics(tp, fp, fn):
    """precision, recall, IoU; NaN where the denominator is empty (eval.py:242-260)."""
    assert not (np.isnan(tp) or np.isnan(fp) or np.isnan(fn)), 'one of the inputs is nan'
    iou = np.nan if (tp == 0 and fp == 0 and fn == 0) else tp / (tp + fp + fn)
    rec = np.nan if tp + fn == 0 else tp / (tp + fn)
    prec = np.nan if tp + fp == 0 else tp / (tp + fp)
    return prec, rec, iou


def get_detection_failures(matched_gts, matched_preds, unique_instance_labels, unique_instance_preds, iou_matrix,
                           precision_matrix, recall_matrix, min_precision_for_pred, min_recall_for_gt):
    """Unmatched predictions / ground truths and, for each, the partner that explains the error (eval.py:35-76):
    a prediction is a commission error only if >= min_precision_for_pred of it lies on labelled trees; an unmatched tree
    whose best prediction recalls >= min_recall_for_gt of it is an under-segmentation, reported with that prediction and the
    other tree the prediction covers."""
    assert (iou_matrix[matched_preds, matched_gts] > 0).sum() == len(matched_preds), 'a zero iou correspondence has been matched'
    non_matched_preds = np.array(list(set(unique_instance_preds) - set(matched_preds))).astype(np.int64)
    non_matched_gts = np.array(list(set(unique_instance_labels) - set(matched_gts))).astype(np.int64)
    pred_gt = [np.nan if precision_matrix[p].sum() < min_precision_for_pred else precision_matrix[p].argmax()
               for p in non_matched_preds]
    gt_pred, gt_other = [], []
    for g in non_matched_gts:
        if recall_matrix[:, g].max() < min_recall_for_gt:
            gt_pred.append(np.nan)
            gt_other.append(np.nan)
            continue
        p = np.argmax(recall_matrix[:, g])
        gt_pred.append(p)
        others = np.delete(np.arange(recall_matrix.shape[1]), g)
        best = recall_matrix[p, others].argmax()
        gt_other.append(np.nan if recall_matrix[p, others][best] < min_recall_for_gt else others[best])
    return non_matched_gts, non_matched_preds, np.array(pred_gt), np.array(gt_pred), np.array(gt_other)


def _partition_table(instance_preds, instance_labels, unique_gts, unique_preds, gt_names, pred_names, intvls, position_of):
    """Shared body of the xy / z partition tables: per matched (pred, gt) pair, metrics inside every interval of the
    normalised coordinate `position_of(ind_positive)` (eval.py:127-226)."""
    import pandas as pd
    cols = {'instance_pred': [], 'instance_label': []}
    spans = [f'intvl{intvls[i]}_{intvls[i + 1]}' for i in range(len(intvls) - 1)]
    for metric in ('prec', 'rec', 'iou'):
        for s in spans:
            cols[f'{metric}_{s}'] = []
    for instance_pred, instance_label in zip(unique_preds, unique_gts):
        cols['instance_pred'].append(pred_names[instance_pred])
        cols['instance_label'].append(gt_names[instance_label])
        ind_pred, ind_pos = instance_preds == instance_pred, instance_labels == instance_label
        rel = position_of(ind_pos)
        for i, s in enumerate(spans):
            sel = (rel >= intvls[i]) & (rel < intvls[i + 1])
            tp, fp, tn, fn = get_eval_components(ind_pred[sel], ind_pos[sel])
            prec, rec, iou = get_segmentation_metrics(tp, fp, fn)
            cols[f'prec_{s}'].append(prec)
            cols[f'rec_{s}'].append(rec)
            cols[f'iou_{s}'].append(iou)
    return pd.DataFrame.from_dict(cols)


def evaluate_no_partition(instance_preds, instance_labels, unique_gts, unique_preds, mapping_to_original_gt_nums,
                          mapping_to_original_pred_nums):
    import pandas as pd
    rows = {'instance_pred': [], 'instance_label': [], 'prec': [], 'rec': [], 'iou': []}
    for instance_pred, instance_label in zip(unique_preds, unique_gts):
        tp, fp, tn, fn = get_eval_components(instance_preds == instance_pred, instance_labels == instance_label)
        prec, rec, iou = get_segmentation_metrics(tp, fp, fn)
        for k, v in zip(rows, (mapping_to_original_pred_nums[instance_pred], mapping_to_original_gt_nums[instance_label], prec, rec, iou)):
            rows[k].append(v)
    return pd.DataFrame.from_dict(rows)


def evaluate_xy_partition(instance_preds, instance_labels, unique_gts, unique_preds, coords, intvls, mapping_to_original_gt_nums,
                          mapping_to_original_pred_nums):
    """Radial partition: distance from the tree position (mean of its points within 0.3 m above its lowest point), in
    units of the tree's 5th-largest distance (eval.py:127-178)."""
    def position_of(ind_pos):
        tree = coords[ind_pos]
        base = np.mean(tree[tree[:, 2] <= np.min(tree[:, 2]) + 0.30], axis=0)[:2]
        dist = np.linalg.norm(coords[:, :2] - base, ord=None, axis=1)
        tree_dist = dist[ind_pos]
        return dist / tree_dist[tree_dist.argsort()[-5]]
    return _partition_table(instance_preds, instance_labels, unique_gts, unique_preds, mapping_to_original_gt_nums,
                            mapping_to_original_pred_nums, intvls, position_of)


def evaluate_z_partition(instance_preds, instance_labels, unique_gts, unique_preds, coords, intvls, mapping_to_original_gt_nums,
                         mapping_to_original_pred_nums):
    """Vertical partition: height above the tree's lowest point in units of its 5th-highest point (eval.py:182-226)."""
    def position_of(ind_pos):
        tree_z = coords[ind_pos][:, 2]
        low = np.min(tree_z)
        top = tree_z[tree_z.argsort()[-5]]
        return (coords - np.array([0, 0, low]))[:, -1] / (top - low)
    return _partition_table(instance_preds, instance_labels, unique_gts, unique_preds, mapping_to_original_gt_nums,
                            mapping_to_original_pred_nums, intvls, position_of)


def evaluate_instance_segmentation(instance_preds, instance_labels, unique_gts, unique_preds, coords, mapping_to_original_gt_nums,
                                   mapping_to_original_pred_nums, xy_partition, z_partition):
    args = (instance_preds, instance_labels, unique_gts, unique_preds)
    maps = (mapping_to_original_gt_nums, mapping_to_original_pred_nums)
    return (evaluate_no_partition(*args, *maps),
            evaluate_xy_partition(*args, coords, xy_partition, *maps) if xy_partition else None,
            evaluate_z_partition(*args, coords, z_partition, *maps) if z_partition else None)


# ---- point-cloud files (data_preparation.py:17-56, util/pipeline.py:339-419) --------------------------------------------
def _laspy():
    try:
        import laspy
        return laspy
    except ImportError as e:
        raise ImportError('LAS / LAZ files need the `laspy` package (not installed); use npy / npz / txt') from e


def load_data(path):
    """[N,4] float array (x, y, z, label) from .npy / .npz / .las / .laz / .txt; unlabeled clouds get label -1."""
    assert path.endswith(('npy', 'npz', 'las', 'laz', 'txt'))
    if path.endswith('npy'):
        data = np.load(path)
    elif path.endswith('npz'):
        data = np.load(path)
        assert 'points' in data
        data = data['points'] if 'labels' not in data else np.hstack((data['points'], data['labels'][:, np.newaxis]))
    elif path.endswith(('.las', '.laz')):
        las = _laspy().read(path)
        points = np.vstack([getattr(las, a) * las.header.scales[i] + las.header.offsets[i] for i, a in enumerate('XYZ')]).T
        if hasattr(las, 'treeID') and hasattr(las, 'classification'):
            tree_id, classes = np.array(las.treeID), np.array(las.classification)
            tree, non_tree = tree_id != 0, np.isin(classes, [1, 2])
            labels = np.ones(len(points))
            labels[tree] = tree_id[tree]
            labels[non_tree] = NON_TREE_CLASS_IN_RAW_DATA
            labels[np.logical_not(tree) & np.logical_not(non_tree)] = INSTANCE_LABEL_IGNORE_IN_RAW_DATA
            data = np.hstack([points, labels[:, np.newaxis]])
        else:
            data = points
    else:
        import pandas as pd
        data = pd.read_csv(path, delimiter=' ').to_numpy()
    assert data.shape[1] in (3, 4)
    if data.shape[1] == 3:
        data = np.hstack([data, INSTANCE_LABEL_IGNORE_IN_RAW_DATA * np.ones(len(data))[:, np.newaxis]])
    return data


def generate_random_color():
    return [random.randint(0, 255) for _ in range(3)]


def save_data(data, save_format, save_name, save_folder, use_offset=True):
    path = osp.join(save_folder, f'{save_name}.{save_format}')
    if save_format in ('las', 'laz'):
        laspy = _laspy()
        assert data.shape[1] == 4
        points, labels = data[:, :3], data[:, 3]
        classification = np.where(labels == 0, 2, 4).astype(labels.dtype)    # terrain / stem (For-Instance convention)
        header = laspy.LasHeader(version='1.2', point_format=3)
        header.offsets = list(points.mean(0)) if use_offset else [0, 0, 0]
        header.scales = [0.001, 0.001, 0.001]
        las = laspy.LasData(header)
        las.x, las.y, las.z = points[:, 0], points[:, 1], points[:, 2]
        las.add_extra_dim(laspy.ExtraBytesParams(name='treeID', type=np.uint32))
        las.treeID = labels
        las.classification = classification
        color_map = {label: generate_random_color() for label in np.unique(labels)}
        colors = np.array([color_map[label] for label in labels], dtype=np.uint16)
        colors[classification == 2] = [0, 0, 0]
        las.red, las.green, las.blue = colors[:, 0], colors[:, 1], colors[:, 2]
        las.write(path)
    elif save_format == 'npy':
        np.save(path, data)
    elif save_format == 'npz':
        np.savez_compressed(path, points=data[:, :3], labels=data[:, 3])
    elif save_format == 'txt':
        np.savetxt(path, data)


def save_treewise(coords, instance_preds, cluster_means_within_hull, insts_not_at_edge, save_format, plot_results_dir,
                  non_trees_label_in_grouping):
    """One file per predicted tree, sorted into completely_inside / trunk_base_inside / trunk_base_outside."""
    coords = coords - np.mean(coords, axis=0)
    dirs = {name: os.path.join(plot_results_dir, name) for name in ('completely_inside', 'trunk_base_inside', 'trunk_base_outside')}
    for d in dirs.values():
        os.makedirs(d, exist_ok=True)
    for i in np.unique(instance_preds):
        pts = coords[instance_preds == i]
        pts = np.hstack([pts, i * np.ones(len(pts))[:, None]])
        if i == non_trees_label_in_grouping:
            save_data(pts, save_format, 'non_trees', plot_results_dir, use_offset=False)
        elif not cluster_means_within_hull[i - 1]:
            save_data(pts, save_format, str(int(i)), dirs['trunk_base_outside'], use_offset=False)
        else:
            where = 'completely_inside' if insts_not_at_edge[i - 1] else 'trunk_base_inside'
            save_data(pts, save_format, str(int(i)), dirs[where], use_offset=False)


# ---- plot outline (util/pipeline.py:211-283; alphashape / shapely / geopandas in the reference) --------------------------
class PlotShape:
    """What `get_hull` / `get_hull_buffer` return here: the exterior ring of the plot outline [V,2] (closed) and, for a
    buffer, the half-width of the band around it.  `get_coords_within_shape` is the only consumer."""

    def __init__(self, ring, buffer=None):
        self.ring, self.buffer = np.asarray(ring, dtype=np.float64), buffer

    def to_pickle(self, path):
        with open(path, 'wb') as f:
            pickle.dump(self, f)


def grid_points(coords, grid_size):
    """First point of every occupied grid_size x grid_size cell, in input order (pipeline.py:226-238)."""
    coords = np.asarray(coords)[:, :2]
    cells = np.floor_divide(coords, grid_size).astype(np.int64)
    _, first = np.unique(cells, axis=0, return_index=True)
    return coords[np.sort(first)]


def _alpha_ring(points, alpha):
    """Exterior ring of the alpha shape: union of the Delaunay triangles with circumradius < 1/alpha (the rule of the
    `alphashape` package); alpha = 0 -> convex hull.  The union must be ONE polygon, as the reference asserts."""
    from scipy.spatial import ConvexHull, Delaunay
    points = np.unique(np.asarray(points, dtype=np.float64), axis=0)
    if alpha <= 0 or len(points) < 4:
        ring = points[ConvexHull(points).vertices]
        return np.vstack([ring, ring[:1]])
    tri = Delaunay(points).simplices
    a, b, c = points[tri[:, 0]], points[tri[:, 1]], points[tri[:, 2]]
    la, lb, lc = np.linalg.norm(b - c, axis=1), np.linalg.norm(a - c, axis=1), np.linalg.norm(a - b, axis=1)
    s = (la + lb + lc) / 2
    area = np.sqrt(np.maximum(s * (s - la) * (s - lb) * (s - lc), 0))
    with np.errstate(divide='ignore', invalid='ignore'):
        keep = la * lb * lc / (4 * area) < 1.0 / alpha
    tri = tri[keep & (area > 0)]
    edges = np.sort(np.vstack([tri[:, [0, 1]], tri[:, [1, 2]], tri[:, [2, 0]]]), axis=1)
    uniq, counts = np.unique(edges, axis=0, return_counts=True)
    boundary = uniq[counts == 1]
    nxt = {}
    for u, v in boundary:
        nxt.setdefault(u, []).append(v)
        nxt.setdefault(v, []).append(u)
    assert boundary.size and all(len(v) == 2 for v in nxt.values()), \
        'failed to calculate concave hull. Set alpha=0 to use convex hull or set outer_remove=~'
    rings, seen = [], set()
    for start in nxt:
        if start in seen:
            continue
        ring, prev, cur = [start], None, start
        seen.add(start)
        while True:
            step = [v for v in nxt[cur] if v != prev]
            prev, cur = cur, (step[0] if step else nxt[cur][0])
            if cur == start:
                break
            ring.append(cur)
            seen.add(cur)
        rings.append(points[ring + [start]])
    area_of = lambda r: abs(np.sum(r[:-1, 0] * r[1:, 1] - r[1:, 0] * r[:-1, 1])) / 2   # noqa: E731
    rings.sort(key=area_of, reverse=True)
    outer = rings[0]
    for r in rings[1:]:      # every other ring must be a hole of the outer one (a second component = MultiPolygon = failure)
        assert _in_ring(r[:1], outer)[0], 'failed to calculate concave hull. Set alpha=0 to use convex hull or set outer_remove=~'
    return outer


def _in_ring(xy, ring):
    """Even-odd point-in-polygon test on the device when one is there (points [n,2], closed ring [V,2])."""
    dev = torch.device('cuda', torch.cuda.current_device()) if torch.cuda.is_available() else torch.device('cpu')
    p = torch.as_tensor(np.asarray(xy, dtype=np.float64), device=dev)
    r = torch.as_tensor(ring, device=dev)
    x0, y0, x1, y1 = r[:-1, 0], r[:-1, 1], r[1:, 0], r[1:, 1]
    inside = torch.zeros(len(p), dtype=torch.bool, device=dev)
    for s in range(0, len(p), 1 << 16):
        px, py = p[s:s + (1 << 16), 0:1], p[s:s + (1 << 16), 1:2]
        crosses = ((y0 > py) != (y1 > py)) & (px < (x1 - x0) * (py - y0) / (y1 - y0) + x0)
        inside[s:s + (1 << 16)] = (crosses.sum(1) % 2) == 1
    return inside.cpu().numpy()


def _near_ring(xy, ring, dist):
    dev = torch.device('cuda', torch.cuda.current_device()) if torch.cuda.is_available() else torch.device('cpu')
    p = torch.as_tensor(np.asarray(xy, dtype=np.float64), device=dev)
    r = torch.as_tensor(ring, device=dev)
    a, d = r[:-1], r[1:] - r[:-1]
    dd = (d * d).sum(1).clamp_min(1e-300)
    near = torch.zeros(len(p), dtype=torch.bool, device=dev)
    for s in range(0, len(p), 1 << 15):
        q = p[s:s + (1 << 15), None, :] - a[None]
        t = ((q * d[None]).sum(2) / dd[None]).clamp(0, 1)
        near[s:s + (1 << 15)] = ((q - t[..., None] * d[None]) ** 2).sum(2).min(1).values < dist * dist
    return near.cpu().numpy()


def _outline(coords, alpha):
    coords = np.asarray(coords)[:, :2]
    mean = np.mean(coords, axis=0, dtype=np.float64)
    return _alpha_ring(grid_points(coords - mean, grid_size=0.25), alpha) + mean


def get_hull(coords, alpha):
    """xy outline of the plot (alpha shape of one point per 0.25 m cell; pipeline.py:258-267)."""
    return PlotShape(_outline(coords, alpha))


def get_hull_buffer(coords, alpha, buffersize):
    """Band of half-width `buffersize` around the outline (pipeline.py:242-254)."""
    return PlotShape(_outline(coords, alpha), buffer=buffersize)


def get_coords_within_shape(coords, shape):
    """Boolean mask of the points whose xy lies within the shape (pipeline.py:211-222)."""
    xy = np.asarray(coords)[:, :2]
    return _near_ring(xy, shape.ring, shape.buffer) if shape.buffer is not None else _in_ring(xy, shape.ring)


def get_cluster_means(coords, labels):
    """Mean coordinate per label, rows in ascending label order (pipeline.py:279-283)."""
    coords, labels = np.asarray(coords), np.asarray(labels)
    uniq, inv = np.unique(labels, return_inverse=True)
    sums = np.zeros((len(uniq), coords.shape[1]))
    np.add.at(sums, inv, coords.astype(np.float64))
    return (sums / np.bincount(inv)[:, None]).astype(coords.dtype)


# ---- file-based tile generation (util/pipeline.py:24-75, data_preparation.py:109-494) ----------------------------------
class SampleGenerator:
    """`SampleGenerator(plot_path, features_path, save_dir, n_neigh_sor, multiplier_sor, rad, npoints_rad)`: tile grid
    cutting of a voxelised plot into `<save_dir>/npz/<plot>_<i>.npz` (+ json meta data), as the pipeline uses it.  The
    denoising filters, `plot_corners` and the random training crops of the reference class are not built."""

    def __init__(self, plot_path, features_path, save_dir, n_neigh_sor, multiplier_sor, rad, npoints_rad):
        data = np.load(plot_path)
        self.points, self.label = data['points'], data['labels']
        self.feats = np.load(features_path)['features']
        self.plot_name = os.path.basename(plot_path)[:-4]
        self.save_dir_data, self.save_dir_meta_data = os.path.join(save_dir, 'npz'), os.path.join(save_dir, 'json')
        os.makedirs(self.save_dir_data, exist_ok=True)
        os.makedirs(self.save_dir_meta_data, exist_ok=True)
        self.n_neigh_sor, self.multiplier_sor, self.rad, self.npoints_rad = n_neigh_sor, multiplier_sor, rad, npoints_rad

    def tile_generate_and_save(self, inner_edge, outer_edge, stride, compressed=False, plot_corners=None, logger=None):
        from . import prepare
        if plot_corners is not None or any(v is not None for v in (self.n_neigh_sor, self.multiplier_sor, self.rad, self.npoints_rad)):
            raise NotImplementedError('plot_corners and the SOR / radius tile filters are not built (unused by the default configs)')
        tiles = prepare.cut_tiles(self.points, self.label, self.feats, inner_edge, outer_edge, stride)
        meta = dict(plot_name=self.plot_name, n_neigh_sor=None, multiplier_sor=None, rad=None, npoints_rad=None,
                    inner_edge=inner_edge, outer_edge=outer_edge)
        save = np.savez_compressed if compressed else np.savez
        for i, tile in enumerate(tiles):
            save(os.path.join(self.save_dir_data, f'{self.plot_name}_{i}.npz'), **tile)
            with open(os.path.join(self.save_dir_meta_data, f'{self.plot_name}_{i}.json'), 'w') as f:
                json.dump(meta, f)

    def __getattr__(self, name):
        raise NotImplementedError(f'SampleGenerator.{name}: random training crops / occupancy grids are offline data '
                                  f'generation outside this build (SURVEY.md §2 row 10)')


class VoxelTrace:
    """What `generate_tiles(return_type='original')` stores as `<plot>_hash_mapping.pkl`: the voxelised coordinates and,
    per voxel, the indices of the original points (CSR) -- the role of the reference's {python hash -> indices} dict."""

    def __init__(self, voxel_coords, offsets, indices):
        self.voxel_coords, self.offsets, self.indices = voxel_coords, offsets, indices


def generate_tiles(cfg, forest_path, logger, return_type='voxelized'):
    """Voxelise the plot, compute verticality, cut the tile grid; same directories / files as the reference
    (`forest_voxelized<v>/<plot>.npz`, `features/<plot>.npz`, `tiles/npz/<plot>_<i>.npz`)."""
    from . import prepare
    plot_name = os.path.basename(forest_path)[:-4]
    base_dir = os.path.dirname(os.path.dirname(forest_path))
    voxelized_dir = osp.join(base_dir, f'forest_voxelized{cfg.voxel_size}')
    features_dir, save_dir = osp.join(base_dir, 'features'), osp.join(base_dir, 'tiles')
    for d in (voxelized_dir, features_dir, save_dir):
        os.makedirs(d, exist_ok=True)
    path_vox = osp.join(voxelized_dir, f'{plot_name}.npz')
    path_idx = osp.join(voxelized_dir, f'{plot_name}_original_idx.pkl')
    path_map = osp.join(voxelized_dir, f'{plot_name}_hash_mapping.pkl')
    logger.info('voxelizing forest...')
    if (not osp.exists(path_vox)) or (return_type == 'original' and not osp.exists(path_idx)):
        data, trace = prepare.voxelize(load_data(forest_path), cfg.voxel_size)
        data = np.round(data.astype(np.float32), 2)
        np.savez_compressed(path_vox, points=data[:, :3], labels=data[:, 3])
        if return_type == 'original':
            with open(path_idx, 'wb') as f:
                pickle.dump(trace, f)
            with open(path_map, 'wb') as f:
                pickle.dump(VoxelTrace(data[:, :3], trace.offsets, trace.indices), f)
    logger.info('calculating features...')
    path_feat = osp.join(features_dir, f'{plot_name}.npz')
    if not osp.exists(path_feat):
        data = load_data(path_vox)
        np.savez_compressed(path_feat, features=prepare.compute_features(data[:, :3].astype(np.float64), cfg.search_radius_features))
    logger.info('getting tiles...')
    cfg.sample_generator.plot_path, cfg.sample_generator.features_path, cfg.sample_generator.save_dir = path_vox, path_feat, save_dir
    SampleGenerator(**cfg.sample_generator).tile_generate_and_save(cfg.inner_edge, cfg.outer_edge, cfg.stride, logger=logger)


def get_hash_values(voxelized_points):
    return [hash(tuple(point)) for point in voxelized_points]


def get_hash_mapping(hash_values, original_idx):
    return {h: original_idx[i] for i, h in enumerate(hash_values)}


def propagate_preds_hash_full(coords, instance_preds, coords_to_return, hash_mapping):
    """Predictions of the voxelised points -> every original point of the voxel (pipeline.py:441-451).  `hash_mapping` is
    the `VoxelTrace` written by this package's `generate_tiles` (join on the device) or, for files written by the reference,
    its {hash(tuple(point)) -> indices} dict (same loop as the reference).  Returns (target_preds, not_yet_propagated)."""
    target = np.empty(coords_to_return.shape[0], np.int64)
    missing = np.ones(coords_to_return.shape[0], bool)
    if isinstance(hash_mapping, VoxelTrace):
        from . import post
        vox_preds, vox_missing = post.propagate_preds_hash_vox(coords, instance_preds, hash_mapping.voxel_coords)
        counts = np.diff(hash_mapping.offsets)
        target[hash_mapping.indices] = np.repeat(vox_preds, counts)
        missing[hash_mapping.indices] = np.repeat(vox_missing, counts)
        return target, missing
    for i, h in enumerate(get_hash_values(np.round(coords, 2))):
        idx = np.array(hash_mapping[h], np.int64)
        target[idx] = instance_preds[i]
        missing[idx] = False
    return target, missing
