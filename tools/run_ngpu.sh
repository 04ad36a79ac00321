# usage: gpurun --gpus N -- 'N=4 bash tools/run_ngpu.sh'
N=${N:-2}; mkdir -p gpurun_out/n$N; O=gpurun_out/n$N
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; tail -3 $O/bench.err; python -c "
import json;d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],json.dumps(d.get('plot')),json.dumps(d.get('train')))"
