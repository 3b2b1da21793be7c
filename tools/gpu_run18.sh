mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'symmetry_finalize' -s 2 -c 1 -o gpurun_out/r18_prof -f python tools/ncu_target.py 888 3 > gpurun_out/r18_ncu.log 2>&1
tail -2 gpurun_out/r18_ncu.log
