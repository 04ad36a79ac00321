mkdir -p gpurun_out/r2g; O=gpurun_out/r2g
timeout 900 python -m pytest tests/test_gpu_halo.py tests/test_gpu_ts.py tests/test_gpu_path.py tests/test_gpu_trained_scale.py -m gpu -x -q > $O/pytest.log 2>&1; tail -5 $O/pytest.log
timeout 300 python tools/profile_layers.py cfg2_2M f16x2 > $O/layers_f16x2.txt 2>&1; sed -n 1,14p $O/layers_f16x2.txt
