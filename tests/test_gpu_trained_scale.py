"""Parity where it is claimed: on BASELINE config 1 (one 20 m tile, ~235 k voxels) with
"trained-like" heads -- the last Linear of both heads fitted so that offsets are metres long and the semantic logits
separate trees (synth.fit_probe_heads; the reference's checkpoint cannot be downloaded) --
  * per-point offsets of the tensor-core modes vs the fp32 oracle (north_star: within 1e-3),
  * instances from the CUDA outputs vs instances from the oracle outputs through the reference's `get_detections`
    (tree_learn/util/eval.py:7-31): same count, every tree matched with IoU >= 0.999."""
import numpy as np
import pytest
import torch

from oracle import cluster_ref, model_ref
from treelearn_b200 import TreeLearn, pipeline, post, synth

pytestmark = pytest.mark.gpu
SHAPE = [500, 500, 1000]
GROUPING = dict(tree_conf_thresh=0.5, tau_vert=0.6, tau_off=4, tau_group=0.15, tau_min=50)
_cache = {}


def _setup():
    """cfg1 tile, probe-fitted state dict, oracle outputs (computed once per session: the oracle takes a few seconds)."""
    if not _cache:
        batch = synth.make_batch([synth.workload('cfg1_200k')])
        sd = model_ref.make_state_dict(channels=32, num_blocks=7, seed=0)
        net = TreeLearn(use_feats=False, use_coords=False, spatial_shape=SHAPE, mode='fp32')
        net.load_state_dict(sd)
        net = net.cuda().eval()
        sd.update(synth.fit_probe_heads(net, batch))
        with torch.no_grad():
            ref = model_ref.forward_ref(sd, batch, spatial_shape=SHAPE)
        _cache.update(batch=batch, sd=sd, ref=ref)
    return _cache['batch'], _cache['sd'], _cache['ref']


def _run(mode, batch, sd, **kw):
    net = TreeLearn(use_feats=False, use_coords=False, spatial_shape=SHAPE, mode=mode, **kw)
    net.load_state_dict(sd)
    net = net.cuda().eval()
    with torch.no_grad():
        out = net(batch, return_loss=False)
    return {k: v.cpu() for k, v in out.items()}


KEEP = 0.9


def _trained_like(batch, ref, out):
    """A random backbone's 32 features cannot say which tree a point belongs to (the probe fit leaves ~2 m rms), so the
    clustering comparison emulates the trained part: the SAME per-point correction  labels - KEEP * oracle output  is
    added to the oracle's and to the CUDA path's outputs.  The oracle side becomes  labels + (1 - KEEP) * oracle
    (predictions 90 % of the way to the labels), and the CUDA side differs from it by exactly its numerical error
    (cuda - oracle), which enters at full size."""
    tree = batch['semantic_labels'] == 0
    c_off = torch.where(tree[:, None], batch['offset_labels'] - KEEP * ref['offset_predictions'], torch.zeros(1))
    want_logits = torch.where(tree, 4.0, -4.0)[:, None] * torch.tensor([1.0, -1.0])
    c_sem = want_logits - KEEP * ref['semantic_prediction_logits']
    return (out['offset_predictions'] + c_off).numpy(), (out['semantic_prediction_logits'] + c_sem).numpy()


def _instances(batch, ref, out):
    coords = batch['coords'].numpy()
    offs, logits = _trained_like(batch, ref, out)
    inst = cluster_ref.get_instances_ref(coords, offs, logits, GROUPING['tree_conf_thresh'],
                                         GROUPING['tau_vert'], GROUPING['tau_off'], GROUPING['tau_group'], GROUPING['tau_min'],
                                         batch['input_feats'].numpy()[:, -1])
    tm = inst != 0
    if (inst[tm] == -1).any() and (inst[tm] != -1).sum() >= 5:
        inst[tm] = cluster_ref.assign_remaining_ref(coords[tm] + offs[tm], inst[tm], -1)
    return inst


def test_probe_heads_give_trained_scale_outputs():
    batch, sd, ref = _setup()
    off = ref['offset_predictions']
    tree = batch['semantic_labels'] == 0
    print('oracle offsets: |max|', float(off.abs().max()), 'p99', float(off.abs().flatten().quantile(0.99)))
    assert off.abs().max() > 3.0                      # metres, like a trained model (tau_off = 4 m)
    acc = ((ref['semantic_prediction_logits'].argmax(1) == 0) == tree).float().mean()
    assert acc > 0.7, acc


@pytest.mark.parametrize('mode,tol', [('f16x2', 1e-3), ('fp32', 1e-3)])
def test_offsets_within_1e3_of_oracle_at_trained_scale(mode, tol):
    batch, sd, ref = _setup()
    out = _run(mode, batch, sd)
    eo = (out['offset_predictions'] - ref['offset_predictions']).abs().max().item()
    el = (out['semantic_prediction_logits'] - ref['semantic_prediction_logits']).abs().max().item()
    print(f'{mode}: max |offset err| {eo:.3e} (tolerance {tol:g}, |offset|max {float(ref["offset_predictions"].abs().max()):.2f} m), '
          f'max |logit err| {el:.3e}')
    assert eo < tol and el < 5 * tol


@pytest.mark.parametrize('k', [1, 2, 3])
def test_mixed_mode_error_by_number_of_two_term_levels(k):
    """Mode 'mixed': two fp16 terms per operand on the first k U-Net levels, one term below.  The error is reported for
    k = 1, 2, 3; the bound asserted is the north_star tolerance for the levels the bench may use (k >= 2) and the measured
    order of magnitude for k = 1."""
    batch, sd, ref = _setup()
    out = _run('mixed', batch, sd, split_levels=k)
    eo = (out['offset_predictions'] - ref['offset_predictions']).abs().max().item()
    el = (out['semantic_prediction_logits'] - ref['semantic_prediction_logits']).abs().max().item()
    rms = (out['offset_predictions'] - ref['offset_predictions']).pow(2).mean().sqrt().item()
    print(f'mixed, {k} two-term level(s): max |offset err| {eo:.3e} (rms {rms:.3e}), max |logit err| {el:.3e}')
    # measured on a B200 (profiles/r02_mixed_mode_error.txt): 1.0e-3 / 2.0e-4 / 6.4e-5 for k = 1 / 2 / 3 at offsets up to 13 m
    assert eo < (5e-4 if k >= 2 else 1e-2)


def test_f16_single_term_error_is_reported_at_trained_scale():
    """Mode f16 (one fp16 term per operand) is the fast mode; at metre-scale offsets its error is of the order of the
    1e-3 budget, which is why bench.py's headline runs f16x2.  The bound asserted here is what the mode does deliver."""
    batch, sd, ref = _setup()
    out = _run('f16', batch, sd)
    eo = (out['offset_predictions'] - ref['offset_predictions']).abs().max().item()
    print(f'f16: max |offset err| {eo:.3e} at |offset|max {float(ref["offset_predictions"].abs().max()):.2f} m')
    assert eo < 2e-2


@pytest.mark.parametrize('mode', ['f16x2', 'f16'])
def test_instances_match_oracle_instances(mode):
    """north_star "instance IoU match on the benchmark tile": the instances clustered from the CUDA outputs equal the ones
    clustered from the oracle's outputs -- same number of trees, every tree matched with IoU >= 0.999 (f16x2) / >= 0.99 (f16).  See _trained_like for how trained heads are emulated."""
    batch, sd, ref = _setup()
    out = _run(mode, batch, sd)
    want, got = _instances(batch, ref, ref), _instances(batch, ref, out)
    n_want, n_got = int(want.max()), int(got.max())
    print(f'{mode}: {n_want} oracle instances, {n_got} from the CUDA outputs')
    assert n_want >= 5, 'degenerate clustering: the probe heads should separate several trees'
    assert n_got == n_want
    tree = want > 0
    mg, mp, iou, _, _ = post.get_detections(want[tree | (got > 0)], got[tree | (got > 0)], 0.5, 0)
    assert len(mg) == n_want
    ious = iou[mp, mg]
    print(f'{mode}: min IoU over matched trees {ious.min():.6f}')
    print(f'{mode}: {int((want != got).sum())} of {len(want)} point labels differ')
    assert ious.min() >= (0.999 if mode == 'f16x2' else 0.99)
