mkdir -p gpurun_out/r2x; O=gpurun_out/r2x
timeout 300 python -m pytest tests/test_gpu_halo.py tests/test_gpu_ts.py -q 2>&1 | tail -8 > $O/pytest_halo.log; cat $O/pytest_halo.log
timeout 200 python tools/profile_layers.py cfg2_2M f16 > $O/layers_f16_halo.txt 2>&1; head -n 14 $O/layers_f16_halo.txt
timeout 200 python tools/profile_layers.py cfg2_2M f16x2 > $O/layers_f16x2_halo.txt 2>&1; head -n 14 $O/layers_f16x2_halo.txt
echo "== TL_HALO_FILL=2 f16"; TL_HALO_FILL=2 timeout 200 python tools/profile_layers.py cfg2_2M f16 2>&1 | sed -n 6,10p
echo "== TL_HALO_FILL=1 f16x2"; TL_HALO_FILL=1 timeout 200 python tools/profile_layers.py cfg2_2M f16x2 2>&1 | sed -n 6,10p
