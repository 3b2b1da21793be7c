mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k plane 2>&1 | tail -4
timeout 300 python tools/bench_stages.py --only plane > gpurun_out/r13_plane.json 2>&1
cat gpurun_out/r13_plane.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r13_launches_plane.csv python tools/bench_stages.py --only plane --reps 1 --skip-cpu --plane-frames 296 > /dev/null 2>&1
