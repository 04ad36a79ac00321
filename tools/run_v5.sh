mkdir -p gpurun_out/r2k; O=gpurun_out/r2k
timeout 600 python -m pytest tests/test_gpu_ts.py -x -q 2>&1 | tail -5 > $O/pytest_ts.log; cat $O/pytest_ts.log
timeout 300 python tools/profile_layers.py cfg2_2M f16 > $O/layers_f16_v5.txt 2>&1; head -n 14 $O/layers_f16_v5.txt
for g in 3 2; do echo "== groups $g"; TL_GRP_GROUPS=$g timeout 200 python tools/profile_layers.py cfg2_2M f16 2>&1 | sed -n 6,9p; done > $O/groups.txt 2>&1; cat $O/groups.txt
timeout 300 python tools/profile_layers.py cfg2_2M f16x2 > $O/layers_f16x2_v5.txt 2>&1; head -n 14 $O/layers_f16x2_v5.txt
TL_LIB=treelearn_b200/libtreelearn_b200_trace.so TL_GRP_DEBUG=32 timeout 200 python tools/trace_ts.py 32 f16 > $O/trace_c32.txt 2>&1; cat $O/trace_c32.txt
for dbg in 1 2 3 7; do echo "== TL_GRP_DEBUG=$dbg"; TL_LIB=treelearn_b200/libtreelearn_b200_trace.so TL_GRP_DEBUG=$dbg timeout 200 python tools/profile_layers.py cfg2_2M f16 2>&1 | sed -n 6,9p; done > $O/ablate.txt 2>&1; cat $O/ablate.txt
