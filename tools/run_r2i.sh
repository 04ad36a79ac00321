mkdir -p gpurun_out/r2i; O=gpurun_out/r2i
timeout 600 python -m pytest tests/test_gpu_halo.py tests/test_gpu_ts.py -m gpu -x -q > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 300 python tools/profile_layers.py cfg2_2M f16x2 > $O/layers_f16x2.txt 2>&1; sed -n 3,14p $O/layers_f16x2.txt
timeout 300 python tools/profile_layers.py cfg2_2M f16 > $O/layers_f16.txt 2>&1; sed -n 3,14p $O/layers_f16.txt
