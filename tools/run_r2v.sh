mkdir -p gpurun_out/r2v; O=gpurun_out/r2v
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -5 $O/pytest_gpu.log
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; tail -3 $O/bench_n1.err; cat $O/bench_n1.json
timeout 200 python tools/profile_step.py cfg2_2M f16x2 > $O/step_f16x2.txt 2>&1; cat $O/step_f16x2.txt | tail -40
