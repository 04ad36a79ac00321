# Round-end validation on the GPU box: GPU tests, smoke, conv traffic capture (then: python tools/summarise_conv_traffic.py <csv> cfg2_2M f16x2 4), bench.
# usage: tools/gpu.sh 2400 "bash tools/run_validation.sh"
mkdir -p gpurun_out/validation; O=gpurun_out/validation
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; tail -4 $O/smoke.txt
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_conv --csv --log-file $O/conv_traffic_f16x2.csv python tools/profile_layers.py cfg2_2M f16x2 > $O/ncu_traffic.log 2>&1; tail -1 $O/ncu_traffic.log | cut -c1-100
python tools/summarise_conv_traffic.py $O/conv_traffic_f16x2.csv cfg2_2M f16x2 4 | cut -c1-300
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; tail -3 $O/bench_n1.err; python -c "
import json;d=json.loads(open('$O/bench_n1.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['fwd_only'],d['e2e'],d['roofline']['frac'],d['roofline']['traffic'],d['mixed_mode']['value'],d['fast_mode']['value'],json.dumps(d['train']['cfg3_4tile_batch'])[:300])"
