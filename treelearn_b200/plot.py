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

from . import pipeline, post, prepare

# tree_learn/dataset/dataset.py:7-10 and tree_learn/util/pipeline.py (grouping labels)
INSTANCE_LABEL_IGNORE_IN_RAW_DATA = -1
NON_TREE_CLASS_IN_RAW_DATA = 0
NON_TREE_CLASS_IN_PYTORCH_DATASET = 1
TREE_CLASS_IN_PYTORCH_DATASET = 0
NON_TREES_LABEL_IN_GROUPING = 0
NOT_ASSIGNED_LABEL_IN_GROUPING = -1
START_NUM_PREDS = 1


def offset_labels(xyz, instance_label, semantic_label):
    """Per-point vector to the tree base = mean of the points within 0.5 m above the (regularised) lowest point of the
    instance (`TreeDataset.getOffset`, dataset.py:121-150).  Returns (offsets f32 [n,3], valid mask)."""
    position = np.ones_like(xyz, dtype=np.float32)
    valid = np.zeros(len(instance_label), dtype=bool)
    for inst in np.unique(instance_label):
        idx = np.where(instance_label == inst)[0]
        if semantic_label[idx[0]] == NON_TREE_CLASS_IN_PYTORCH_DATASET:
            continue
        z = xyz[idx, 2]
        low = np.partition(z, 10)[3] if len(z) > 11 else z.min()
        near_base = xyz[idx][z <= low + 0.5]
        if len(near_base) > 0:
            position[idx] = np.mean(near_base, axis=0)
            valid[idx] = True
        else:
            position[idx] = np.array([0, 0, 0])
    return position - xyz, valid


def tile_sample(tile, inner_square_edge_length):
    """One tile dict of `prepare.cut_tiles` -> the tensors `TreeDataset.__getitem__` returns in test mode."""
    xyz, inst = tile['points'], tile['instance_label']
    sem = np.where(inst == NON_TREE_CLASS_IN_RAW_DATA, NON_TREE_CLASS_IN_PYTORCH_DATASET,
                   TREE_CLASS_IN_PYTORCH_DATASET).astype(np.float64)
    center = np.ones_like(xyz) * tile['center']
    off, off_valid = offset_labels(xyz, inst, sem)
    inner = np.linalg.norm(xyz[:, :-1], ord=np.inf, axis=1) <= (inner_square_edge_length / 2)
    keep = inst != INSTANCE_LABEL_IGNORE_IN_RAW_DATA
    return dict(coords=torch.from_numpy(xyz), input_feats=torch.from_numpy(tile['feat']),
                instance_labels=torch.from_numpy(inst), semantic_labels=torch.from_numpy(sem),
                offset_labels=torch.from_numpy(off), centers=torch.from_numpy(center),
                masks_inner=torch.from_numpy(inner), masks_sem=torch.from_numpy(inner & keep),
                masks_off=torch.from_numpy(inner & keep & (sem != NON_TREE_CLASS_IN_PYTORCH_DATASET) & off_valid))


def tiles_to_batches(tiles, inner_square_edge_length, batch_size=1):
    """Generator of model input dicts (`TreeDataset.collate_fn`, dataset.py:176-226) over the tiles, in order."""
    as_type = dict(coords=torch.float32, input_feats=torch.float32, semantic_labels=torch.long, instance_labels=torch.long,
                   masks_inner=torch.bool, masks_off=torch.bool, masks_sem=torch.bool, offset_labels=torch.float32,
                   centers=torch.float32)
    for start in range(0, len(tiles), batch_size):
        samples = [tile_sample(t, inner_square_edge_length) for t in tiles[start:start + batch_size]]
        batch = {k: torch.cat([s[k] for s in samples], 0).to(dt) for k, dt in as_type.items()}
        batch['batch_ids'] = torch.cat([torch.full((len(s['coords']),), b, dtype=torch.long) for b, s in enumerate(samples)])
        batch['batch_size'] = len(samples)
        yield batch


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
