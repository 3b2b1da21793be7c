mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "plane" 2>&1 | tail -6
timeout 600 python tools/bench_stages.py --only plane > gpurun_out/r11_plane.json 2>&1
cat gpurun_out/r11_plane.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r11_launches_plane.csv python tools/bench_stages.py --only plane --reps 1 --skip-cpu > /dev/null 2>&1
tail -14 gpurun_out/r11_launches_plane.csv | cut -d, -f5,15 | cut -c1-80
