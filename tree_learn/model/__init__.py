from treelearn_b200.model import TreeLearn  # noqa: F401  (reference: tree_learn/model/__init__.py:1)
