mkdir -p gpurun_out
timeout 600 python tools/variant_probe.py 1024 > gpurun_out/r8_variants.log 2>&1
cat gpurun_out/r8_variants.log
