mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k regex:'points_kernel|integral_kernel|distmap_kernel|normals_kernel|ccl_|candidates_kernel|fit_kernel|model_image|refine_kernel' -s 13 -c 13 -o gpurun_out/rp_prof -f python tools/bench_stages.py --only plane --reps 1 --skip-cpu --plane-frames 296 > gpurun_out/rp_ncu.log 2>&1
tail -2 gpurun_out/rp_ncu.log
