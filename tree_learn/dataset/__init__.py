"""`from tree_learn.dataset import TreeDataset` (reference tree_learn/dataset/__init__.py:1)."""
from treelearn_b200.dataset import TreeDataset  # noqa: F401
