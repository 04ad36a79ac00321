mkdir -p gpurun_out/r2a; O=gpurun_out/r2a
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -6 $O/pytest_gpu.log
timeout 300 python tools/profile_train.py 4 tf32 3 2>&1 | tail -2 > $O/train_cfg3.txt; cat $O/train_cfg3.txt
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; tail -3 $O/bench_n1.err; python -c "
import json;d=json.loads(open('$O/bench_n1.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['fwd_only'],d['e2e']['value'],d['roofline']['frac'],json.dumps(d.get('train')))"
