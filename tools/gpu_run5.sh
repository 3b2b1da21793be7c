mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 600 python tools/e2e_probe.py 1024 6 > gpurun_out/r5_e2e_probe.log 2>&1
cat gpurun_out/r5_e2e_probe.log
