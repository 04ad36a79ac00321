"""`TreeDataset` -- the tile dataset the reference's tools construct (`tree_learn.dataset.TreeDataset`,
reference tree_learn/dataset/dataset.py:13-226): same constructor, `__getitem__` tuple, `collate_fn` batch dict (the
model's input contract) and the same sequence of `np.random` draws in training mode, so a seeded run sees
the same augmented crops as the reference.  Host-side numpy / torch, as `north_star` prescribes for the dataset.

The in-memory whole-plot path (`treelearn_b200.plot`) builds the same samples without files through `sample_from_arrays`.
"""
import math
import os

import numpy as np
import torch
from torch.utils.data import Dataset

INSTANCE_LABEL_IGNORE_IN_RAW_DATA = -1      # dataset.py:7-10
NON_TREE_CLASS_IN_RAW_DATA = 0
NON_TREE_CLASS_IN_PYTORCH_DATASET = 1
TREE_CLASS_IN_PYTORCH_DATASET = 0

_BATCH_DTYPES = dict(coords=torch.float32, input_feats=torch.float32, semantic_labels=torch.long, instance_labels=torch.long,
                     masks_inner=torch.bool, masks_off=torch.bool, masks_sem=torch.bool, offset_labels=torch.float32,
                     centers=torch.float32)
_SAMPLE_ORDER = ('coords', 'input_feats', 'instance_labels', 'semantic_labels', 'offset_labels', 'centers', 'masks_inner',
                 'masks_off', 'masks_sem')


def offset_labels(xyz, instance_label, semantic_label):
    """Per-point vector to the tree base = mean of the instance's points within 0.5 m above its (regularised) lowest
    point (`getOffset`, dataset.py:111-140).  Returns (offsets [n,3] in xyz's promoted dtype, valid mask)."""
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


def augment(xyz, cfg, prob=0.5, prob_point_jitter=0.25):
    """Training augmentation (`transform_train` + `dataAugment`, dataset.py:92-108,143-164): optional per-point jitter,
    then ONE random 3x3 matrix (anisotropic scale, matrix jitter, x flip, z rotation).  Draw order = the reference's."""
    if cfg['point_jitter'] == True and np.random.random() <= prob_point_jitter:   # noqa: E712  (the reference's test)
        xyz += np.clip(0.1 * np.random.randn(xyz.shape[0], 3), -0.2, 0.2)
    m = np.eye(3)
    if cfg['scaled'] and np.random.rand() < prob:
        m = m * np.concatenate([np.random.uniform(0.8, 1.2, 2), np.random.uniform(0.95, 1.05, 1)])
    if cfg['jitter'] and np.random.rand() < prob:
        m += np.random.randn(3, 3) * 0.1
    if cfg['flip'] and np.random.rand() < prob:
        m[0][0] *= np.random.randint(0, 2) * 2 - 1
    if cfg['rot'] and np.random.rand() < prob:
        theta = np.random.rand() * 2 * math.pi
        m = np.matmul(m, [[math.cos(theta), math.sin(theta), 0], [-math.sin(theta), math.cos(theta), 0], [0, 0, 1]])
    return np.matmul(xyz, m)


def sample_from_arrays(xyz, feat, inst, center, inner_square_edge_length, augmentations=None):
    """The nine tensors of one sample, keyed by the batch-dict names.  `center` None = training (dummy ones)."""
    sem = np.empty(len(inst))
    sem[inst == NON_TREE_CLASS_IN_RAW_DATA] = NON_TREE_CLASS_IN_PYTORCH_DATASET
    sem[inst != NON_TREE_CLASS_IN_RAW_DATA] = TREE_CLASS_IN_PYTORCH_DATASET
    centers = np.ones_like(xyz) if center is None else np.ones_like(xyz) * center
    if augmentations is not None:
        xyz = augment(xyz, augmentations)
    off, off_valid = offset_labels(xyz, inst, sem)
    inner = np.linalg.norm(xyz[:, :-1], ord=np.inf, axis=1) <= (inner_square_edge_length / 2)
    keep = np.logical_not(inst == INSTANCE_LABEL_IGNORE_IN_RAW_DATA)
    arrays = dict(coords=xyz, input_feats=feat, instance_labels=inst, semantic_labels=sem, offset_labels=off, centers=centers,
                  masks_inner=inner, masks_sem=inner & keep,
                  masks_off=inner & keep & (sem != NON_TREE_CLASS_IN_PYTORCH_DATASET) & off_valid)
    return {k: torch.from_numpy(v) for k, v in arrays.items()}


def collate_samples(samples):
    """List of sample dicts -> the model's batch dict (`collate_fn`, dataset.py:167-226)."""
    assert len(samples) > 0, 'empty batch'
    batch = {k: torch.cat([s[k] for s in samples], 0).to(dt) for k, dt in _BATCH_DTYPES.items()}
    batch['batch_ids'] = torch.cat([torch.full((len(s['coords']),), b, dtype=torch.long) for b, s in enumerate(samples)])
    batch['batch_size'] = len(samples)
    return batch


class TreeDataset(Dataset):
    def __init__(self, data_root, inner_square_edge_length, training, logger, data_augmentations=None):
        self.data_paths = [os.path.join(data_root, path) for path in os.listdir(data_root)]
        self.inner_square_edge_length = inner_square_edge_length
        self.logger = logger
        self.training = training
        self.data_augmentations = data_augmentations
        self.logger.info(f'Load {"train" if training else "test"} dataset: {len(self.data_paths)} scans')

    def __len__(self):
        return len(self.data_paths)

    def __getitem__(self, index):
        data = np.load(self.data_paths[index])
        s = sample_from_arrays(data['points'], data['feat'], data['instance_label'], None if self.training else data['center'],
                               self.inner_square_edge_length, self.data_augmentations if self.training else None)
        return tuple(s[k] for k in _SAMPLE_ORDER)

    def collate_fn(self, batch):
        return collate_samples([dict(zip(_SAMPLE_ORDER, sample)) for sample in batch])
