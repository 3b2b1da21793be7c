mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'extract_kernel|symmetry_finalize' -s 4 -c 2 -o gpurun_out/rf_prof -f python tools/ncu_target.py 888 3 > gpurun_out/rf_ncu.log 2>&1
tail -2 gpurun_out/rf_ncu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/rf_launches_bench.csv python bench.py --steps 2 --warmup 3 --frames 512 --cpu-sample 8 > gpurun_out/rf_bench_under_ncu.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/rf_bench_n1.json 2> gpurun_out/rf_bench.err
cut -c1-200 gpurun_out/rf_bench_n1.json
timeout 600 python tools/bench_stages.py > gpurun_out/rf_stages.json 2>> gpurun_out/rf_bench.err
cut -c1-160 gpurun_out/rf_stages.json
