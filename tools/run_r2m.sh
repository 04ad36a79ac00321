mkdir -p gpurun_out/r2m; O=gpurun_out/r2m
timeout 900 python -m pytest tests/test_gpu_halo.py tests/test_gpu_path.py tests/test_gpu_properties.py -m gpu -x -q > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 300 python tools/profile_step.py cfg2_2M f16x2 > $O/step_f16x2.txt 2>&1; tail -11 $O/step_f16x2.txt
