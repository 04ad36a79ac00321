"""`from tree_learn.util import ...` -- hot-path functions come from the CUDA-backed implementation; the reference's
CPU pre/post-processing helpers (tile generation, hulls, LAS I/O, evaluation, plotting: SURVEY.md §2 rows 5, 10-16)
are out of this build's scope and raise a clear error when touched."""
from treelearn_b200.pipeline import (assign_remaining_points_nearest_neighbor, ensemble, get_instances,  # noqa: F401
                                      get_pointwise_preds, group_dbscan, group_hdbscan, make_labels_consecutive)
from treelearn_b200.post import get_detections, propagate_preds, propagate_preds_hash_vox  # noqa: F401
from treelearn_b200.prepare import compute_features, voxelize  # noqa: F401
from treelearn_b200.train_util import (build_dataloader, build_optimizer, checkpoint_save, cuda_cast,  # noqa: F401
                                        is_multiple, load_checkpoint, point_wise_loss)

_OUT_OF_SCOPE = {'generate_tiles', 'get_coords_within_shape', 'get_hull_buffer', 'get_hull', 'get_cluster_means',
                 'save_treewise', 'load_data', 'save_data', 'propagate_preds_hash_full', 'get_config', 'get_args_and_cfg', 'munch_to_dict', 'get_root_logger',
                 'init_train_logger', 'build_cosine_scheduler', 'get_eval_components', 'SampleGenerator'}


def __getattr__(name):
    if name in _OUT_OF_SCOPE:
        raise NotImplementedError(f'tree_learn.util.{name} is CPU pre/post-processing outside the per-tile hot path '
                                  f'(SURVEY.md §8 scope table); use the reference implementation for it')
    raise AttributeError(name)
