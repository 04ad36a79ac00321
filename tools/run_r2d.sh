mkdir -p gpurun_out/r2d; O=gpurun_out/r2d
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -x -q > $O/pytest_train.log 2>&1; tail -5 $O/pytest_train.log
timeout 300 python tools/profile_train.py 4 tf32 4 2>&1 | tail -3 > $O/train_cfg3.txt; cat $O/train_cfg3.txt
timeout 300 python tools/profile_train.py 2 tf32 4 2>&1 | tail -2 > $O/train_2tiles.txt; cat $O/train_2tiles.txt
