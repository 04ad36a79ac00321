mkdir -p gpurun_out/r2q; O=gpurun_out/r2q
timeout 300 python -m pytest tests/test_gpu_ts.py -x -q 2>&1 | tail -5 > $O/pytest_ts.log; cat $O/pytest_ts.log
timeout 200 python tools/profile_layers.py cfg2_2M f16 > $O/layers_f16_v7.txt 2>&1; head -n 14 $O/layers_f16_v7.txt
timeout 200 python tools/profile_layers.py cfg2_2M f16x2 > $O/layers_f16x2_v7.txt 2>&1; head -n 14 $O/layers_f16x2_v7.txt
for cfg in "1" "3"; do echo "== fill $cfg"; TL_GRP_FILL=$cfg timeout 200 python tools/profile_layers.py cfg2_2M f16x2 2>&1 | sed -n 6,13p; done > $O/fill_x2.txt 2>&1; cat $O/fill_x2.txt
timeout 600 python -m pytest tests/test_gpu_trained_scale.py -q -s 2>&1 | grep -v "^$" | tail -40 > $O/pytest_trained.log; cat $O/pytest_trained.log
