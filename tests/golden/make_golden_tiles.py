"""Golden vectors for tile cutting: the tiles the REFERENCE's own `SampleGenerator.tile_generate_and_save`
(/root/reference/tree_learn/util/data_preparation.py:333-494; `.cuda()` made a no-op on this CPU-only box) writes for
seeded voxelised plots.  The voxel down-sampling / verticality halves of that row call open3d / jakteristics, which are
absent here: they stay "parity unpinned" (oracle/prepare_ref.py).  Run in the build container:

    python tests/golden/make_golden_tiles.py        -> tests/golden/tiles_small.npz
"""
import logging
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference  # noqa: E402

from oracle import prepare_ref  # noqa: E402
from treelearn_b200 import synth  # noqa: E402


def plot_case(seed, edge, shift, inner_edge, outer_edge, stride):
    """A voxelised plot as generate_tiles stores it: fp32 points rounded to 2 decimals, fp32 labels, fp32 verticality."""
    f = synth.synth_forest(edge=edge, height=8.0, n_trees=6, seed=seed, ground_density=40.0)
    pts = np.round((f['coords'].astype(np.float64) + np.array(shift)).astype(np.float32), 2)
    return pts, f['inst'].astype(np.float32), f['feat'].astype(np.float32).reshape(-1, 1), inner_edge, outer_edge, stride


def main():
    import_reference()
    import tree_learn.util.data_preparation as ref_prep
    import tree_learn.dataset.dataset as ref_dataset
    assert ref_prep.__file__.startswith('/root/reference')
    logger = logging.getLogger('golden')
    cases = {'a': plot_case(1, 30.0, (120.37, -45.81, 3.0), 8, 13.5, 0.5),          # the configured default edges
             'b': plot_case(2, 18.0, (-7.5, 1003.21, 0.0), 5, 6.0, 1.0),            # no overlap
             'c': plot_case(3, 24.0, (0.0, 0.0, 0.0), 6, 4.0, 0.25)}                 # dense overlap
    out = {}
    for name, (pts, lab, feat, inner_edge, outer_edge, stride) in cases.items():
        with tempfile.TemporaryDirectory() as tmp:
            np.savez(os.path.join(tmp, 'plot.npz'), points=pts, labels=lab)
            np.savez(os.path.join(tmp, 'feat.npz'), features=feat)
            gen = ref_prep.SampleGenerator(os.path.join(tmp, 'plot.npz'), os.path.join(tmp, 'feat.npz'),
                                           os.path.join(tmp, 'tiles'), None, None, None, None)
            gen.tile_generate_and_save(inner_edge, outer_edge, stride, logger=logger)
            files = sorted(os.listdir(os.path.join(tmp, 'tiles', 'npz')), key=lambda s: int(s[:-4].split('_')[-1]))
            tiles = [dict(np.load(os.path.join(tmp, 'tiles', 'npz', f))) for f in files]
            # the model input the reference's TreeDataset (test mode) + collate_fn build from the first two tiles
            ds = ref_dataset.TreeDataset(os.path.join(tmp, 'tiles', 'npz'), inner_edge, False, logger)
            ds.data_paths = [os.path.join(tmp, 'tiles', 'npz', f) for f in files[:2]]
            batch = ds.collate_fn([ds[0], ds[1]]) if name != 'a' else {}
            for key, val in batch.items():
                out[f'{name}:batch:{key}'] = val.numpy() if hasattr(val, 'numpy') else np.asarray(val)
        ora = prepare_ref.cut_tiles_ref(pts, lab, feat, inner_edge, outer_edge, stride)
        assert len(ora) == len(tiles), (name, len(ora), len(tiles))
        for t, (a, b) in enumerate(zip(tiles, ora)):
            for key in ('points', 'feat', 'instance_label', 'center'):
                assert a[key].dtype == b[key].dtype and np.array_equal(a[key], b[key]), (name, t, key)
        print(name, len(pts), 'points ->', len(tiles), 'tiles,', sum(len(t['points']) for t in tiles), 'rows')
        out[f'{name}:points'], out[f'{name}:labels'], out[f'{name}:feats'] = pts, lab, feat
        out[f'{name}:cfg'] = np.array([inner_edge, outer_edge, stride], dtype=np.float64)
        out[f'{name}:n_tiles'] = len(tiles)
        out[f'{name}:rows'] = np.array([len(t['points']) for t in tiles])
        out[f'{name}:centers'] = np.array([t['center'] for t in tiles])
        # the full tiles are reproducible from the inputs; keep a checksum per tile plus two complete tiles
        out[f'{name}:sum_points'] = np.array([t['points'].astype(np.float64).sum(axis=0) for t in tiles])
        out[f'{name}:sum_labels'] = np.array([t['instance_label'].astype(np.int64).sum() for t in tiles])
        for t in (0, len(tiles) - 1):
            for key in ('points', 'feat', 'instance_label'):
                out[f'{name}:tile{t}:{key}'] = tiles[t][key]
    path = os.path.join(HERE, 'tiles_small.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, f'{os.path.getsize(path) / 1024:.0f} KiB')


if __name__ == '__main__':
    main()
