mkdir -p gpurun_out/r2x; O=gpurun_out/r2x
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -x -q -s > $O/pytest_train.log 2>&1; tail -15 $O/pytest_train.log
for w in 0 1; do echo "== TL_WGRAD_TC=$w 4 tiles"; TL_WGRAD_TC=$w timeout 300 python tools/profile_train.py 4 tf32 3 2>&1 | tail -4; done > $O/train_ab.txt 2>&1
echo "== TL_WGRAD_TC=1 2 tiles" >> $O/train_ab.txt; timeout 300 python tools/profile_train.py 2 tf32 4 2>&1 | tail -5 >> $O/train_ab.txt
cat $O/train_ab.txt
