"""`from tree_learn.util import ...` -- the flat namespace of the reference (tree_learn/util/__init__.py:3-9).
Hot-path functions come from the CUDA-backed implementation (pipeline / post / prepare), the host-side helpers
(config, logging, scheduler, metrics, I/O, hulls, file-based tile generation) from `treelearn_b200.host_util`.
Everything the reference's tools import resolves at import time; what depends on a package that is not installed
(LAS files -> laspy) raises when CALLED."""
from treelearn_b200.pipeline import (assign_remaining_points_nearest_neighbor, ensemble, get_instances,  # noqa: F401
                                      get_pointwise_preds, group_dbscan, group_hdbscan, make_labels_consecutive)
from treelearn_b200.post import get_detections, propagate_preds, propagate_preds_hash_vox  # noqa: F401
from treelearn_b200.prepare import compute_features, voxelize  # noqa: F401
from treelearn_b200.train_util import (build_dataloader, build_optimizer, checkpoint_save, cuda_cast,  # noqa: F401
                                        is_multiple, load_checkpoint, point_wise_loss)
from treelearn_b200.host_util import (SampleGenerator, SummaryWriter, build_cosine_scheduler,  # noqa: F401
                                       evaluate_instance_segmentation, evaluate_no_partition, evaluate_xy_partition,
                                       evaluate_z_partition, generate_random_color, generate_tiles, get_args,
                                       get_args_and_cfg, get_cluster_means, get_config, get_coords_within_shape,
                                       get_detection_failures, get_eval_components, get_hash_mapping, get_hash_values,
                                       get_hull, get_hull_buffer, get_root_logger, get_segmentation_metrics, grid_points,
                                       init_train_logger, load_data, load_yaml_file, modify_default_cfg, munch_to_dict,
                                       propagate_preds_hash_full, save_data, save_treewise)
