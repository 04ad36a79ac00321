"""CPU tests: the oracle against the committed golden vectors (produced by the reference's own code,
tests/golden/make_golden.py) and against its own dense cross-check forms."""
import os

import numpy as np
import torch

from oracle import cluster_ref, model_ref, spconv_ref as sp

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def load_model_fixture():
    g = np.load(os.path.join(GOLD, 'model_small.npz'))
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith('sd:')}
    batch = {k[6:]: torch.from_numpy(g[k]) for k in g.files if k.startswith('batch:')}
    batch['batch_size'] = int(batch['batch_size'])
    out = {k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith('out:')}
    return g, sd, batch, out


def test_oracle_model_matches_reference_code_golden():
    g, sd, batch, out = load_model_fixture()
    with torch.no_grad():
        ora = model_ref.forward_ref(sd, batch, use_coords=False, use_feats=True, spatial_shape=[500, 500, 1000])
        loss, ld = model_ref.loss_ref(ora, batch)
    for k, v in out.items():
        assert torch.allclose(ora[k], v, atol=2e-5, rtol=1e-5), k
    assert abs(loss.item() - float(g['loss'])) < 1e-4
    assert abs(ld['semantic_loss'].item() - float(g['semantic_loss'])) < 1e-4
    assert abs(ld['offset_loss'].item() - float(g['offset_loss'])) < 1e-4


def test_oracle_training_grads_match_reference_code_golden():
    g, sd, batch, _ = load_model_fixture()
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k else v) for k, v in sd.items()}
    out = model_ref.forward_ref(sd, batch, use_coords=False, use_feats=True, spatial_shape=[500, 500, 1000], training=True)
    loss, _ = model_ref.loss_ref(out, batch)
    assert abs(loss.item() - float(g['train_loss'])) < 1e-3
    loss.backward()
    for k in [k[5:] for k in g.files if k.startswith('grad:')]:
        ref = torch.from_numpy(g['grad:' + k])
        assert torch.allclose(sd[k].grad, ref, atol=1e-4 + 1e-3 * ref.abs().max().item()), k


def _random_sparse(n=400, shape=(12, 10, 14), batch=2, c=5, seed=0):
    rng = np.random.default_rng(seed)
    idx = np.unique(np.stack([rng.integers(0, batch, n), rng.integers(0, shape[0], n), rng.integers(0, shape[1], n),
                              rng.integers(0, shape[2], n)], 1), axis=0)
    rng.shuffle(idx)
    return torch.from_numpy(idx.astype(np.int32)), torch.from_numpy(rng.normal(size=(len(idx), c)).astype(np.float32))


def test_rulebook_convs_equal_dense_forms():
    shape, batch = [12, 10, 14], 2
    idx, feats = _random_sparse(shape=shape, batch=batch)
    torch.manual_seed(0)
    x = sp.SparseConvTensor(feats, idx, shape, batch)
    subm = sp.SubMConv3d(5, 7, 3, padding=1, bias=False, indice_key='s')
    down = sp.SparseConv3d(7, 6, 2, stride=2, bias=False, indice_key='d')
    up = sp.SparseInverseConv3d(6, 4, 2, bias=False, indice_key='d')
    with torch.no_grad():
        a = subm(x)
        assert torch.allclose(a.features, sp.dense_subm(feats, idx, shape, batch, subm.weight), atol=1e-5)
        b = down(a)
        assert torch.allclose(b.features, sp.dense_strided(a.features, idx, shape, batch, down.weight, b.indices), atol=1e-5)
        c = up(b)
        assert torch.equal(c.indices, idx)
        assert torch.allclose(c.features, sp.dense_inverse(b.features, b.indices, b.spatial_shape, batch, up.weight,
                                                           idx, shape), atol=1e-5)


def test_strided_odd_edge_and_reach_zero():
    idx = torch.tensor([[0, 4, 0, 0], [0, 3, 1, 1], [0, 0, 0, 0]], dtype=torch.int32)
    out_idx, out_shape, in_row, kappa, out_row = sp.strided_pairs(idx.numpy(), [5, 4, 4])
    assert out_shape == [2, 2, 2]
    assert sorted(in_row.tolist()) == [1, 2]          # x=4 -> q=2 >= 2 is dropped (odd edge)
    try:
        sp.strided_pairs(idx.numpy(), [1, 4, 4])
        assert False
    except ValueError as e:
        assert 'reach zero!!!' in str(e)


def test_cluster_oracle_matches_reference_golden():
    g = np.load(os.path.join(GOLD, 'cluster_small.npz'))
    names = ['coords', 'semantic_scores', 'semantic_labels', 'offset_predictions', 'offset_labels', 'instance_labels',
             'feats', 'input_feats']
    e_in = {n: g['ens_in:' + n] for n in names}
    out = cluster_ref.ensemble_ref(**e_in)
    for n, a in zip(names, out):
        ref = g['ens_out:' + n]
        assert a.dtype == ref.dtype and a.shape == ref.shape, n
        assert np.allclose(a, ref, rtol=1e-5, atol=1e-6), n
    assert np.array_equal(out[0], g['ens_out:coords'])
    inst = cluster_ref.get_instances_ref(out[0], out[3], out[1], 0.5, 0.6, 4, 0.15, 50, out[7][:, -1])
    assert np.array_equal(inst, g['instances'])
    tm = inst != 0
    assigned = cluster_ref.assign_remaining_ref(out[0][tm] + out[3][tm], inst[tm], -1)
    assert np.array_equal(assigned, g['assigned'])
    assert np.array_equal(cluster_ref.radius_components(g['p2'], 0.15), g['p2_raw'])
    assert np.array_equal(cluster_ref.group_dbscan_ref(g['p2'], 0.15, 20, -1, 1), g['p2_group'])


def test_loss_matches_reference_golden():
    g = np.load(os.path.join(GOLD, 'cluster_small.npz'))
    out = {'semantic_prediction_logits': torch.from_numpy(g['ens_out:semantic_scores']),
           'offset_predictions': torch.from_numpy(g['ens_out:offset_predictions'])}
    batch = {'masks_sem': torch.from_numpy(g['loss_masks_sem']), 'masks_off': torch.from_numpy(g['loss_masks_off']),
             'semantic_labels': torch.from_numpy(g['loss_labels']), 'offset_labels': torch.from_numpy(g['loss_offset_labels'])}
    _, ld = model_ref.loss_ref(out, batch)
    assert abs(ld['semantic_loss'].item() / 50 - float(g['loss_semantic'])) < 1e-5
    assert abs(ld['offset_loss'].item() - float(g['loss_offset'])) < 1e-5
