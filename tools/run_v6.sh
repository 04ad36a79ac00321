mkdir -p gpurun_out/r2m; O=gpurun_out/r2m
timeout 600 python -m pytest tests/test_gpu_ts.py -x -q 2>&1 | tail -3 > $O/pytest_ts.log; cat $O/pytest_ts.log
for cfg in "1 224 0" "0 224 0" "1 160 0" "1 128 0" "1 128 3" "1 128 2" "1 100 0" "1 96 2"; do set -- $cfg
  echo "== CA=$1 SMEM_KB=$2 GROUPS=$3"; TL_GRP_CA=$1 TL_GRP_SMEM_KB=$2 TL_GRP_GROUPS=$3 timeout 200 python tools/profile_layers.py cfg2_2M f16 2>&1 | sed -n 6,10p; done > $O/sweep.txt 2>&1; cat $O/sweep.txt
