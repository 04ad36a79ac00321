"""Whole-plot inference in memory (treelearn_b200/plot.py): raw points -> tiles -> network -> merge -> instances ->
predictions back on the voxelised / original points, checked stage by stage against the oracle chain."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import cluster_ref, model_ref, post_ref, prepare_ref
from treelearn_b200 import plot, synth

GROUPING = SimpleNamespace(tree_conf_thresh=0.5, tau_vert=0.3, tau_off=4, tau_group=0.15, tau_min=10, use_hdbscan=False)
TILES = dict(inner_edge=4, outer_edge=3.0, stride=0.5)
MODEL = dict(channels=8, num_blocks=3, use_feats=True, use_coords=False, spatial_shape=[500, 500, 1000])


def raw_plot(seed=4, n=40000):
    f = synth.synth_forest(edge=9.0, height=6.0, n_trees=4, seed=seed, ground_density=60.0)
    rng = np.random.default_rng(seed)
    pick = rng.integers(0, len(f['coords']), n)
    return f['coords'][pick].astype(np.float64) + rng.normal(0, 0.03, (n, 3)) + np.array([5120.4, -310.9, 40.0])


def oracle_merge(sd, plot_xyz_label, feats):
    """Tiles -> oracle forward per tile -> inner rows -> oracle overlap merge, from the prepared plot."""
    tiles = prepare_ref.cut_tiles_ref(plot_xyz_label[:, :3], plot_xyz_label[:, 3], feats, **TILES)
    keep = {k: [] for k in ('coords', 'logits', 'sem', 'off', 'off_lab', 'inst', 'feats', 'in_feats')}
    for batch in plot.tiles_to_batches(tiles, TILES['inner_edge'], batch_size=2):
        with torch.no_grad():
            out = model_ref.forward_ref(sd, batch, use_coords=False, use_feats=True, spatial_shape=MODEL['spatial_shape'])
        inner = batch['masks_inner']
        keep['coords'].append((batch['coords'] + batch['centers'])[inner])
        keep['logits'].append(out['semantic_prediction_logits'][inner])
        keep['off'].append(out['offset_predictions'][inner])
        keep['feats'].append(out['backbone_feats'][inner])
        for k, name in (('sem', 'semantic_labels'), ('off_lab', 'offset_labels'), ('inst', 'instance_labels'),
                        ('in_feats', 'input_feats')):
            keep[k].append(batch[name][inner])
    cat = {k: torch.cat(v).numpy() for k, v in keep.items()}
    return len(tiles), cluster_ref.ensemble_ref(cat['coords'], cat['logits'], cat['sem'], cat['off'], cat['off_lab'],
                                                cat['inst'], cat['feats'], cat['in_feats'])


def oracle_instances(coords, offsets, logits, verticality):
    tree_mask = (torch.from_numpy(logits).float().softmax(dim=-1)[:, 0] >= GROUPING.tree_conf_thresh).numpy()
    inst = cluster_ref.get_instances_ref(coords, offsets, logits, GROUPING.tree_conf_thresh, GROUPING.tau_vert,
                                         GROUPING.tau_off, GROUPING.tau_group, GROUPING.tau_min, verticality,
                                         tree_mask=tree_mask)
    tm = inst != 0
    inst[tm] = cluster_ref.assign_remaining_ref(coords[tm] + offsets[tm], inst[tm], -1)
    return inst


def oracle_propagate(coords, preds, target):
    out, missing = post_ref.propagate_preds_hash_vox_ref(coords, preds, target)
    if missing.any():
        out[missing] = post_ref.propagate_preds_ref(coords, preds, target[missing], 5)
    return out, missing


def test_oracle_chain_runs_on_an_oracle_prepared_plot():
    """CPU: the checker used by the GPU test below is itself exercised end to end (and stays fast)."""
    data = raw_plot(n=6000)
    centred = data - data.mean(0)
    down, _ = prepare_ref.voxelize_ref(np.hstack([centred, -np.ones((len(data), 1))]), 0.1)
    plot_arr = np.round(down.astype(np.float32), 2)
    feats = prepare_ref.compute_features_ref(plot_arr[:, :3].astype(np.float64), 0.6)
    sd = model_ref.make_state_dict(channels=8, num_blocks=3, seed=7)
    n_tiles, merged = oracle_merge(sd, plot_arr, feats)
    coords, logits, _, offsets, _, _, _, in_feats = merged
    assert n_tiles > 4 and len(coords) <= len(plot_arr)
    inst = oracle_instances(coords, offsets, logits, in_feats[:, -1])
    out, missing = oracle_propagate(coords, inst, plot_arr[:, :3])
    assert out.shape == (len(plot_arr),) and missing.mean() < 0.05


@pytest.mark.gpu
@pytest.mark.parametrize('return_type', ['voxelized', 'original'])
def test_segment_points_stage_by_stage_vs_oracle(return_type):
    from treelearn_b200 import TreeLearn
    data = raw_plot()
    sd = model_ref.make_state_dict(channels=8, num_blocks=3, seed=7)
    net = TreeLearn(**MODEL)
    net.load_state_dict(sd)
    net = net.cuda().eval()
    res = plot.segment_points(net, data, SimpleNamespace(voxel_size=0.1), GROUPING, voxel_size=0.1,
                              search_radius_features=0.6, return_type=return_type, batch_size=2, **TILES)
    plot_arr, feats = res['plot'], res['features']
    # 1. tiles + network + overlap merge against the oracle chain on the same prepared plot
    n_tiles, merged = oracle_merge(sd, plot_arr, feats)
    coords, logits, _, offsets, _, _, _, in_feats = merged
    assert res['n_tiles'] == n_tiles > 4
    assert np.array_equal(res['voxel_coords'], coords)
    assert np.abs(res['offset_predictions'] - offsets).max() < 1e-3
    assert np.abs(res['semantic_prediction_logits'] - logits).max() < 1e-3
    assert np.allclose(res['input_feats'], in_feats, rtol=0, atol=1e-6)
    # 2. instances from the product's own merged predictions: label for label
    want = oracle_instances(res['voxel_coords'], res['offset_predictions'], res['semantic_prediction_logits'],
                            res['input_feats'][:, -1])
    assert np.array_equal(res['voxel_instance_preds'], want)
    assert want.max() >= 1 and (want == 0).any()                      # the fixture yields trees and non-tree points
    # 3. predictions back on the plot / the input points
    vox, missing = oracle_propagate(res['voxel_coords'], want, plot_arr[:, :3])
    if return_type == 'voxelized':
        assert np.array_equal(res['instance_preds'], vox)
        assert np.allclose(res['coords'], plot_arr[:, :3].astype(np.float64) + data.mean(0), atol=1e-9)
    else:
        trace = res['trace']
        assert np.array_equal(res['coords'], data) or np.allclose(res['coords'], data, rtol=0, atol=1e-9)
        ok = np.repeat(~missing, np.diff(trace.offsets))
        assert np.array_equal(res['instance_preds'][trace.indices][ok], np.repeat(vox, np.diff(trace.offsets))[ok])
        assert res['instance_preds'].shape == (len(data),) and res['instance_preds'].min() >= 0


@pytest.mark.gpu
def test_segment_plot_several_tiles_per_forward_equals_one_tile_per_forward():
    """dist.segment_plot runs a rank's tiles in chunks of several tiles per network forward (one batch element per tile);
    the merged plot and its instance labels are the same as with one tile per forward (the reference's batch size 1)."""
    from treelearn_b200 import TreeLearn
    from treelearn_b200 import dist as tdist
    g = SimpleNamespace(tree_conf_thresh=0.5, tau_vert=0.6, tau_off=4, tau_group=0.15, tau_min=50, use_hdbscan=False)
    forest = synth.synth_forest(edge=24.0, n_trees=30, seed=7)
    xyz = torch.from_numpy(forest['coords'])
    tiles = []
    for cx in (-6.0, 0.0, 6.0):
        for cy in (-6.0, 0.0, 6.0):                      # 3 x 3 tiles of 14 m with an 8 m inner square: real overlaps
            sel = (((xyz[:, 0] - cx).abs() < 7.0) & ((xyz[:, 1] - cy).abs() < 7.0)).numpy()
            t = {k: (v[sel] if hasattr(v, 'shape') and len(v) == len(xyz) else v) for k, v in forest.items()}
            t['coords'] = (t['coords'] - [cx, cy, 0.0]).astype('float32')
            t['centre'] = (forest['centre'] + [cx, cy, 0.0]).astype('float32')
            tiles.append(synth.make_batch([t], inner_edge=8.0))
    torch.manual_seed(0)
    net = synth.randomize_bn_stats(TreeLearn(use_feats=False, use_coords=False, spatial_shape=[500, 500, 1000],
                                             mode='f16x2')).cuda().eval()
    c1, l1, n1 = tdist.segment_plot(net, tiles, g, points_per_forward=1)              # one tile per forward
    c4, l4, n4 = tdist.segment_plot(net, tiles, g, points_per_forward=10 ** 9)        # all nine tiles in one forward
    c2, l2, n2 = tdist.segment_plot(net, tiles, g, points_per_forward=int(2.5 * tiles[0]['coords'].shape[0]))
    assert torch.equal(c1, c4) and torch.equal(c1, c2)
    assert torch.equal(l1, l4) and torch.equal(l1, l2) and int(n1) == int(n4) == int(n2)
