mkdir -p gpurun_out/r2r; O=gpurun_out/r2r
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -25 > $O/pytest_gpu.log; cat $O/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -5 $O/smoke.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; tail -3 $O/bench.err; cat $O/bench.json
for m in f16x2 f16; do
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_conv --csv --log-file $O/conv_traffic_$m.csv python tools/profile_layers.py cfg2_2M $m > $O/ncu_layers_$m.log 2>&1; tail -2 $O/ncu_layers_$m.log
done
