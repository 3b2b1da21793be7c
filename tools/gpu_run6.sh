mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r6_launches_e2e.csv python tools/e2e_probe.py 1024 1 4 8 > gpurun_out/r6_probe.log 2>&1
tail -3 gpurun_out/r6_probe.log
grep -c . gpurun_out/r6_launches_e2e.csv
