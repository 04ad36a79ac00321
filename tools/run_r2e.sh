mkdir -p gpurun_out/r2e; O=gpurun_out/r2e
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; tail -3 $O/bench_n2.err; python -c "
import json;d=json.loads(open('$O/bench_n2.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],json.dumps(d.get('plot')),json.dumps(d.get('train')))"
