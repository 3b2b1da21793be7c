mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'extract_kernel|symmetry_finalize' -s 4 -c 2 -o gpurun_out/r7_prof -f python tools/ncu_target.py 888 3 > gpurun_out/r7_ncu.log 2>&1
tail -3 gpurun_out/r7_ncu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r7_bench.json 2> gpurun_out/r7_bench.err
cat gpurun_out/r7_bench.json; tail -3 gpurun_out/r7_bench.err
