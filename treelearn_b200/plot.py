"""Whole-plot inference in memory: the stages of the reference's `run_treelearn_pipeline`
(tools/pipeline/pipeline.py:20-200) chained on the GPU without the .npz / .pkl files between them.

    raw plot -> centre (:40-43) -> voxelize + verticality + tiles (generate_tiles, util/pipeline.py:24-75)
             -> TreeDataset test-mode samples + collate (tree_learn/dataset/dataset.py:35-226)
             -> get_pointwise_preds -> ensemble -> get_instances -> assign_remaining_points_nearest_neighbor (:66-94)
             -> predictions back on the voxelised or the original points (:160-184) -> de-centre (:187)

Not built (third-party geometry): alpha-shape hulls / outer-point removal (`shape_cfg`), per-tree
files, LAS output.
"""
import numpy as np
import torch

from . import dataset, pipeline, post, prepare
from .dataset import offset_labels  # noqa: F401  (kept under its old name)

# tree_learn/dataset/dataset.py:7-10 and tree_learn/util/pipeline.py (grouping labels)
INSTANCE_LABEL_IGNORE_IN_RAW_DATA = -1
NON_TREE_CLASS_IN_RAW_DATA = 0
NON_TREE_CLASS_IN_PYTORCH_DATASET = 1
TREE_CLASS_IN_PYTORCH_DATASET = 0
NON_TREES_LABEL_IN_GROUPING = 0
NOT_ASSIGNED_LABEL_IN_GROUPING = -1
START_NUM_PREDS = 1


def tile_sample(tile, inner_square_edge_length):
    """One tile dict of `prepare.cut_tiles` -> the tensors `TreeDataset.__getitem__` returns in test mode."""
    return dataset.sample_from_arrays(tile['points'], tile['feat'], tile['instance_label'], tile['center'],
                                      inner_square_edge_length)


def tiles_to_batches(tiles, inner_square_edge_length, batch_size=1):
    """Generator of model input dicts (`TreeDataset.collate_fn`, dataset.py:176-226) over the tiles, in order."""
    for start in range(0, len(tiles), batch_size):
        yield dataset.collate_samples([tile_sample(t, inner_square_edge_length) for t in tiles[start:start + batch_size]])


def segment_points(model, data, model_cfg, grouping_cfg, voxel_size=0.1, search_radius_features=0.6, inner_edge=8,
                   outer_edge=13.5, stride=0.5, return_type='voxelized', batch_size=1, logger=None):
    """data [N, 3 or 4] (x, y, z[, instance label]) -> dict with
         'coords' [P,3] f64 (voxelised plot or, return_type='original', the input points) and 'instance_preds' [P] i64
         (0 = not a tree, 1.. = trees), plus the merged pointwise results under the reference's names."""
    if return_type not in ('voxelized', 'original'):
        raise ValueError(f'return_type {return_type!r}: expected "voxelized" or "original"')
    data = np.asarray(data)
    xyz = data[:, :3].astype(np.float64)
    xyz_mean = np.mean(xyz, 0).astype(np.float64)
    labels = data[:, 3:4].astype(np.float64) if data.shape[1] >= 4 else INSTANCE_LABEL_IGNORE_IN_RAW_DATA * np.ones((len(xyz), 1))
    plot, feats, tiles, trace = prepare.prepare_tiles(np.hstack([xyz - xyz_mean, labels]), voxel_size, search_radius_features,
                                                      inner_edge, outer_edge, stride)
    pointwise = pipeline.get_pointwise_preds(model, tiles_to_batches(tiles, inner_edge, batch_size), model_cfg, logger)
    merged = pipeline.ensemble(*[pointwise[i] for i in (4, 0, 1, 2, 3, 5, 6, 7)])
    coords, logits, sem_labels, offsets, off_labels, inst_labels, backbone_feats, input_feats = merged
    preds = pipeline.get_instances(coords, offsets, logits, grouping_cfg, input_feats[:, -1], TREE_CLASS_IN_PYTORCH_DATASET,
                                   NON_TREES_LABEL_IN_GROUPING, NOT_ASSIGNED_LABEL_IN_GROUPING, START_NUM_PREDS)
    initial = np.copy(preds)
    tree = preds != NON_TREES_LABEL_IN_GROUPING
    preds[tree] = pipeline.assign_remaining_points_nearest_neighbor(coords[tree] + offsets[tree], preds[tree],
                                                                   NOT_ASSIGNED_LABEL_IN_GROUPING)
    # back onto the voxelised plot by exact coordinates; `original`: every input point takes the prediction of its voxel
    # (the reference's hash_mapping, here the trace); what found no partner is filled in from its 5 nearest predicted
    # voxels (tools/pipeline/pipeline.py:160-184)
    vox_preds, vox_missing = post.propagate_preds_hash_vox(coords, preds, plot[:, :3])
    if return_type == 'voxelized':
        out_coords, out_preds, missing = plot[:, :3], vox_preds, vox_missing
    else:
        counts = np.diff(trace.offsets)
        out_coords = xyz - xyz_mean
        out_preds, missing = np.empty(len(xyz), dtype=np.int64), np.empty(len(xyz), dtype=bool)
        out_preds[trace.indices], missing[trace.indices] = np.repeat(vox_preds, counts), np.repeat(vox_missing, counts)
    if missing.any():
        out_preds[missing] = post.propagate_preds(coords, preds, out_coords[missing], n_neighbors=5)
    out_coords = out_coords.astype(np.float64) + xyz_mean
    return {'coords': out_coords, 'instance_preds': out_preds, 'voxel_coords': coords, 'voxel_instance_preds': preds,
            'instance_preds_after_initial_clustering': initial, 'offset_predictions': offsets,
            'semantic_prediction_logits': logits, 'input_feats': input_feats, 'n_tiles': len(tiles), 'plot': plot,
            'features': feats, 'trace': trace}
