"""treelearn_b200 -- B200-native (sm_100a) per-tile hot path of ecker-lab/TreeLearn.

Public surface mirrors the reference's names for this path:
  TreeLearn                                  <- tree_learn.model.TreeLearn
  get_pointwise_preds, ensemble, get_instances, group_dbscan, make_labels_consecutive,
  assign_remaining_points_nearest_neighbor   <- tree_learn.util.pipeline
  point_wise_loss, cuda_cast                 <- tree_learn.util.train
"""
from .model import TreeLearn, point_wise_loss  # noqa: F401

__all__ = ['TreeLearn', 'point_wise_loss']
