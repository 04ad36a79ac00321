mkdir -p gpurun_out/r2w; O=gpurun_out/r2w
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -x -q -k "autocast" -s > $O/pytest_autocast.log 2>&1; tail -5 $O/pytest_autocast.log
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; tail -3 $O/bench_n1.err; python -c "
import json;d=json.loads(open('$O/bench_n1.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],json.dumps(d.get('train')))"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-plot --no-train --no-cluster > $O/ncu_bench.log 2>&1; tail -2 $O/ncu_bench.log | cut -c1-300
python tools/summarise_launches.py $O/launches.csv > $O/launches_summary.txt; head -40 $O/launches_summary.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_conv_halo -s 2 -c 2 -o $O/halo_f16x2_c32 python tools/profile_layers.py cfg2_2M f16x2 > $O/ncu_f16x2.log 2>&1; tail -3 $O/ncu_f16x2.log
ls -la $O
