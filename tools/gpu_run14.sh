mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --cpu-sample 64 > gpurun_out/r14_bench.json 2> gpurun_out/r14_bench.err
cut -c1-220 gpurun_out/r14_bench.json; grep -o '"e2e": {[^}]*}' gpurun_out/r14_bench.json; tail -3 gpurun_out/r14_bench.err
timeout 600 python tools/bench_stages.py --only config3 --skip-cpu > gpurun_out/r14_config3.json 2>&1; cat gpurun_out/r14_config3.json
