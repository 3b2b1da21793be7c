set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r3_tests.log
cat gpurun_out/r3_tests.log
timeout 300 python tools/sym_counters.py > gpurun_out/r3_counters.log 2>&1
cat gpurun_out/r3_counters.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r3_bench.json 2> gpurun_out/r3_bench.err
cat gpurun_out/r3_bench.json; tail -3 gpurun_out/r3_bench.err
timeout 600 python tools/e2e_probe.py 1024 6 > gpurun_out/r3_e2e_probe.log 2>&1
cat gpurun_out/r3_e2e_probe.log
