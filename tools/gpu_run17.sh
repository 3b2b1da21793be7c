mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/extract_phases.py 444 > gpurun_out/r17_phases.log 2>&1; tail -6 gpurun_out/r17_phases.log
timeout 600 python bench.py --steps 10 --warmup 3 --cpu-sample 64 > gpurun_out/r17_bench.json 2> gpurun_out/r17_bench.err
cut -c1-220 gpurun_out/r17_bench.json; grep -o '"e2e": {[^}]*}' gpurun_out/r17_bench.json | cut -c1-80; grep -o '"kernel_ms": {[^}]*}' gpurun_out/r17_bench.json
