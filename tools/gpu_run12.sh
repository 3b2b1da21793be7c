mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'refine_kernel|distmap_kernel|normals_kernel|ccl_strip' -s 10 -c 4 -o gpurun_out/r12_prof -f python tools/bench_stages.py --only plane --reps 1 --skip-cpu > gpurun_out/r12_ncu.log 2>&1
tail -2 gpurun_out/r12_ncu.log
