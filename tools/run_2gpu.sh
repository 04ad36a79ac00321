mkdir -p gpurun_out/r2s; O=gpurun_out/r2s
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dist_check.py > $O/dist_check_2gpu.log 2>&1; tail -15 $O/dist_check_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; tail -3 $O/bench_n2.err; cat $O/bench_n2.json
