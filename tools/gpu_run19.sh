mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 --cpu-sample 64 > gpurun_out/r19_bench.json 2> gpurun_out/r19_bench.err
cut -c1-200 gpurun_out/r19_bench.json; grep -o '"e2e": {[^}]*}' gpurun_out/r19_bench.json | cut -c1-60,330-420; grep -o '"kernel_ms": {[^}]*}' gpurun_out/r19_bench.json
