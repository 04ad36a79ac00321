"""Generate the committed golden fixtures by running the REFERENCE's own code in this container.

Run from the repo root (only here: /root/reference does not exist on the GPU box):
    python tests/golden/make_golden.py

What is pinned and how:
  * model_small.npz -- the reference's OWN tree_learn/model/{tree_learn,blocks}.py executed
    verbatim, with `oracle.spconv_ref` injected as `spconv` (spconv itself is not installable:
    the spconv boundary stays "parity unpinned") and seeded weights in the reference's
    state_dict layout.  Pins module structure, key names, BN/ReLU/residual/concat order, heads.
  * cluster_small.npz -- the reference's OWN `ensemble`, `get_instances` (DBSCAN branch),
    `group_dbscan`, `assign_remaining_points_nearest_neighbor` and `point_wise_loss`
    (sklearn / pandas), tree_learn/util/pipeline.py:113-296, util/train.py:145-166.
Optional third-party imports of the reference that are absent here are stubbed with MagicMock.
"""
import os
import sys
from unittest.mock import MagicMock

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = '/root/reference'

from oracle import spconv_ref, model_ref, cluster_ref  # noqa: E402
from treelearn_b200 import synth  # noqa: E402


def import_reference():
    for name in ['geopandas', 'alphashape', 'laspy', 'shapely', 'shapely.geometry', 'open3d', 'jakteristics',
                 'tensorboardX', 'timm', 'timm.scheduler', 'munch', 'plotly', 'plotly.express',
                 'plotly.graph_objects', 'torchvision', 'torchvision.datasets', 'torchvision.datasets.utils']:
        sys.modules.setdefault(name, MagicMock())
    spconv_ref.install_as_spconv()
    # this repo ships a drop-in package also called `tree_learn`; make sure the REFERENCE one wins here
    for k in [k for k in sys.modules if k == 'tree_learn' or k.startswith('tree_learn.')]:
        del sys.modules[k]
    # (the reference's `tree_learn` has no __init__.py: a regular package of the same name anywhere on sys.path would
    # shadow it, so path entries holding this repo's drop-in package are hidden while the reference is imported)
    saved = list(sys.path)
    sys.path[:] = [REF] + [p for p in saved if not os.path.isfile(os.path.join(p or '.', 'tree_learn', '__init__.py'))]
    torch.Tensor.cuda = lambda self, *a, **k: self     # cuda_cast (util/train.py:28-43) on a CPU box
    import tree_learn.model as ref_model
    import tree_learn.util.pipeline as ref_pipeline
    import tree_learn.util.train as ref_train
    sys.path[:] = [REF] + saved
    assert ref_model.__file__.startswith(REF), ref_model.__file__
    return ref_model, ref_pipeline, ref_train


def model_fixture(ref_model, out_path):
    cfg = dict(channels=8, num_blocks=3, use_feats=True, use_coords=False, spatial_shape=[500, 500, 1000])
    tile_a = synth.synth_forest(edge=4.0, n_trees=2, seed=11, ground_density=120.0)
    tile_b = synth.synth_forest(edge=3.0, n_trees=1, seed=12, ground_density=150.0)
    batch = synth.make_batch([tile_a, tile_b], inner_edge=2.0)
    sd = model_ref.make_state_dict(channels=8, num_blocks=3, seed=5)
    net = ref_model.TreeLearn(**cfg)
    missing, unexpected = net.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    assert list(net.state_dict().keys()) == list(sd.keys()), 'state_dict layout differs from the reference'
    net.eval()
    with torch.no_grad():
        ref_out = net(dict(batch), return_loss=False)
        ref_loss, ref_ld = net(dict(batch), return_loss=True)
    with torch.no_grad():
        ora = model_ref.forward_ref(sd, batch, use_coords=False, use_feats=True, spatial_shape=cfg['spatial_shape'])
        ora_loss, _ = model_ref.loss_ref(ora, batch)
    for k in ref_out:
        err = (ref_out[k] - ora[k]).abs().max().item()
        print(f'oracle vs reference-code  {k}: max abs err {err:.3e}')
        assert err < 2e-5, k
    assert abs(ref_loss.item() - ora_loss.item()) < 1e-4
    # training-mode forward/backward of the reference code (BN batch stats, autograd)
    net.train()
    tl, _ = net(dict(batch), return_loss=True)
    tl.backward()
    grads = {k: p.grad.clone() for k, p in net.named_parameters()}
    gsel = ['input_conv.0.weight', 'unet.blocks.block0.conv_branch.2.weight', 'unet.conv.2.weight',
            'unet.u.u.blocks.block1.conv_branch.5.weight', 'unet.deconv.2.weight',
            'unet.blocks_tail.block0.i_branch.0.weight', 'unet.blocks_tail.block0.conv_branch.0.weight',
            'offset_linear.3.weight', 'semantic_linear.0.bias']
    np.savez_compressed(
        out_path, cfg_channels=8, cfg_num_blocks=3,
        **{'sd:' + k: v.numpy() for k, v in sd.items()},
        **{'batch:' + k: (v.numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in batch.items()},
        **{'out:' + k: v.numpy() for k, v in ref_out.items()},
        loss=ref_loss.item(), semantic_loss=ref_ld['semantic_loss'].item(), offset_loss=ref_ld['offset_loss'].item(),
        train_loss=tl.item(), **{'grad:' + k: grads[k].numpy() for k in gsel})
    print('wrote', out_path, os.path.getsize(out_path) // 1024, 'KiB')


def cluster_fixture(ref_pipeline, ref_train, out_path):
    rng = np.random.default_rng(7)
    tile = synth.synth_forest(edge=8.0, n_trees=4, seed=21, ground_density=150.0)
    coords = tile['coords'] + np.float32(100.0)                 # re-globalised like util/pipeline.py:99
    n = len(coords)
    tree = tile['inst'] > 0
    offs = np.where(tree[:, None], tile['base'] - tile['coords'], 0).astype(np.float32)
    offs += rng.normal(0, 0.05, offs.shape).astype(np.float32)
    logits = np.stack([np.where(tree, 2.0, -2.0), np.where(tree, -2.0, 2.0)], 1).astype(np.float32)
    logits += rng.normal(0, 1.5, logits.shape).astype(np.float32)
    vert = tile['feat']
    # ---- ensemble: every point seen in 1-3 "tiles" with jittered predictions
    rep = rng.integers(1, 4, n)
    src = np.repeat(np.arange(n), rep)
    perm = rng.permutation(len(src))
    src = src[perm]
    e_in = dict(coords=coords[src] + rng.choice([0, 1e-4, -1e-4], (len(src), 3)).astype(np.float32),
                semantic_scores=logits[src] + rng.normal(0, 0.1, (len(src), 2)).astype(np.float32),
                semantic_labels=(~tree[src]).astype(np.int64),
                offset_predictions=offs[src] + rng.normal(0, 0.02, (len(src), 3)).astype(np.float32),
                offset_labels=offs[src], instance_labels=tile['inst'][src],
                feats=rng.normal(size=(len(src), 4)).astype(np.float32), input_feats=vert[src].reshape(-1, 1))
    e_out = ref_pipeline.ensemble(**e_in)
    names = ['coords', 'semantic_scores', 'semantic_labels', 'offset_predictions', 'offset_labels',
             'instance_labels', 'feats', 'input_feats']
    o_out = cluster_ref.ensemble_ref(**e_in)
    for nm, a, b in zip(names, e_out, o_out):
        assert a.shape == b.shape and a.dtype == b.dtype, nm
        assert np.allclose(a, b, rtol=1e-5, atol=1e-6), nm
    assert np.array_equal(e_out[0], o_out[0])
    # ---- get_instances (DBSCAN branch) + remaining-point assignment on the merged cloud
    from types import SimpleNamespace
    g = SimpleNamespace(tree_conf_thresh=0.5, tau_vert=0.6, tau_off=4, tau_group=0.15, tau_min=50, use_hdbscan=False)
    m_coords, m_logits, _, m_offs, _, _, _, m_in = e_out
    inst = ref_pipeline.get_instances(m_coords, m_offs, m_logits, g, m_in[:, -1], 0, 0, -1, 1)
    o_inst = cluster_ref.get_instances_ref(m_coords, m_offs, m_logits, 0.5, 0.6, 4, 0.15, 50, m_in[:, -1])
    assert np.array_equal(inst, o_inst), 'oracle get_instances differs from the reference'
    tm = inst != 0
    shifted = m_coords[tm] + m_offs[tm]
    assigned = ref_pipeline.assign_remaining_points_nearest_neighbor(shifted, inst[tm], -1)
    o_assigned = cluster_ref.assign_remaining_ref(shifted, inst[tm], -1)
    assert np.array_equal(assigned, o_assigned), 'oracle kNN assignment differs from the reference'
    # raw DBSCAN labels on a denser 2-D set (label numbering by lowest index)
    p2 = rng.normal(size=(4000, 2)).astype(np.float32) * np.float32(0.8)
    from sklearn.cluster import DBSCAN
    raw = DBSCAN(eps=0.15, min_samples=2).fit(p2).labels_
    assert np.array_equal(raw, cluster_ref.radius_components(p2, 0.15))
    grp = ref_pipeline.group_dbscan(p2.copy(), 0.15, 20, -1, 1)
    assert np.array_equal(grp, cluster_ref.group_dbscan_ref(p2, 0.15, 20, -1, 1))
    # loss
    lg = torch.from_numpy(m_logits)
    of = torch.from_numpy(m_offs)
    lab = torch.from_numpy((rng.uniform(size=len(lg)) < 0.5).astype(np.int64))
    olab = torch.from_numpy(rng.normal(size=m_offs.shape).astype(np.float32))
    ms = torch.from_numpy(rng.uniform(size=len(lg)) < 0.7)
    mo = torch.from_numpy(rng.uniform(size=len(lg)) < 0.3)
    sl, ol = ref_train.point_wise_loss(lg, of, ms, mo, lab, olab)
    np.savez_compressed(
        out_path, **{'ens_in:' + k: v for k, v in e_in.items()}, **{'ens_out:' + nm: a for nm, a in zip(names, e_out)},
        instances=inst, tree_mask=tm, assigned=assigned, p2=p2, p2_raw=raw, p2_group=grp,
        loss_labels=lab.numpy(), loss_offset_labels=olab.numpy(), loss_masks_sem=ms.numpy(), loss_masks_off=mo.numpy(),
        loss_semantic=sl.item(), loss_offset=ol.item())
    print('clusters:', inst.max(), 'unassigned before kNN:', int((inst == -1).sum()),
          'wrote', out_path, os.path.getsize(out_path) // 1024, 'KiB')


if __name__ == '__main__':
    torch.manual_seed(0)
    ref_model, ref_pipeline, ref_train = import_reference()
    here = os.path.dirname(os.path.abspath(__file__))
    model_fixture(ref_model, os.path.join(here, 'model_small.npz'))
    cluster_fixture(ref_pipeline, ref_train, os.path.join(here, 'cluster_small.npz'))
