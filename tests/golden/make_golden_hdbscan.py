"""Golden vectors for the HDBSCAN row: outputs of the REFERENCE's own `group_hdbscan`
(/root/reference/tree_learn/util/pipeline.py:184-191 -> sklearn.cluster.HDBSCAN, sklearn 1.9.0 in this image) and of
`get_instances(..., use_hdbscan=True)` on seeded inputs, plus sklearn's raw labels.  Run in the build container:

    python tests/golden/make_golden_hdbscan.py        -> tests/golden/hdbscan_small.npz
"""
import os
import sys
import warnings
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference  # noqa: E402


def tree_bases(seed, n_trees, noise, edge):
    """Offset-shifted trunk points look like this: dense blobs of very different sizes at the tree bases + stragglers."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(0, edge, (n_trees, 2))
    pts = [c[i] + rng.normal(0, rng.uniform(0.04, 0.2), (int(rng.integers(30, 420)), 2)) for i in range(n_trees)]
    pts.append(rng.uniform(0, edge, (noise, 2)))
    return np.concatenate(pts).astype(np.float32)


def main():
    warnings.simplefilter('ignore')
    _, ref_pipeline, _ = import_reference()
    from sklearn.cluster import HDBSCAN
    out = {}
    cases = {'a': (tree_bases(1, 30, 1200, 25.0), 50),
             'b': (np.round(tree_bases(2, 18, 600, 18.0), 2).astype(np.float32), 50),      # 1 cm grid: ties + duplicates
             'c': (tree_bases(3, 8, 150, 10.0), 20),
             'd': (np.random.default_rng(4).uniform(0, 6, (900, 2)).astype(np.float32), 50)}   # no real structure
    for name, (pts, mcs) in cases.items():
        raw = HDBSCAN(min_cluster_size=mcs).fit_predict(pts)
        grouped = ref_pipeline.group_hdbscan(pts.copy(), mcs, -1, 1)
        out[f'{name}:points'], out[f'{name}:mcs'] = pts, mcs
        out[f'{name}:sklearn_labels'], out[f'{name}:group_hdbscan'] = raw.astype(np.int64), np.asarray(grouped).astype(np.int64)
        print(name, len(pts), 'points,', int(raw.max()) + 1, 'clusters,', int((raw == -1).sum()), 'noise')
    # get_instances with the default clusterer switch, on the merged-plot fixture of cluster_small.npz
    g = np.load(os.path.join(HERE, 'cluster_small.npz'))
    cfg = SimpleNamespace(tree_conf_thresh=0.5, tau_vert=0.6, tau_off=4, tau_group=0.15, tau_min=50, use_hdbscan=True)
    inst = ref_pipeline.get_instances(g['ens_out:coords'], g['ens_out:offset_predictions'], g['ens_out:semantic_scores'], cfg,
                                      g['ens_out:input_feats'][:, -1], 0, 0, -1, 1)
    out['instances_hdbscan'] = np.asarray(inst).astype(np.int64)
    print('get_instances(use_hdbscan=True):', len(inst), 'points,', int(inst.max()), 'instances')
    path = os.path.join(HERE, 'hdbscan_small.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB')


if __name__ == '__main__':
    main()
