mkdir -p gpurun_out/r2k; O=gpurun_out/r2k
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
timeout 300 python tools/profile_step.py cfg2_2M f16x2 > $O/step_f16x2.txt 2>&1; tail -12 $O/step_f16x2.txt
