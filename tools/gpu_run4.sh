mkdir -p gpurun_out
timeout 600 python tools/e2e_probe.py 1024 6 > gpurun_out/r4_e2e_probe.log 2>&1
cat gpurun_out/r4_e2e_probe.log
timeout 900 python -m pytest tests -m gpu -x -q -k "extraction or pixels or full_batch" 2>&1 | tail -5
