mkdir -p gpurun_out/r2z; O=gpurun_out/r2z
(
timeout 120 python tools/debug_wgrad.py 32 32
TL_WG_SBO=1024 TL_WG_LBO_A=512 TL_WG_LBO_B=512 timeout 120 python tools/debug_wgrad.py 32 32
timeout 120 python tools/debug_wgrad.py 64 96
) > $O/debug_wgrad.txt 2>&1; cat $O/debug_wgrad.txt
timeout 600 python -m pytest tests/test_gpu_halo.py -m gpu -x -q -s > $O/pytest_halo.log 2>&1; tail -4 $O/pytest_halo.log
timeout 300 python tools/profile_layers.py cfg2_2M f16x2 > $O/layers_f16x2.txt 2>&1; sed -n 1,8p $O/layers_f16x2.txt
