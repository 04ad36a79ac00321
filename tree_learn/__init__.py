"""Drop-in alias package: the names tools/pipeline and tools/training import from the reference's `tree_learn`
resolve to the B200-native implementation in `treelearn_b200`."""
