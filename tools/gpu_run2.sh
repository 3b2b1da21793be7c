set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'extract_kernel|symmetry_finalize' -s 4 -c 2 -o gpurun_out/r2_prof -f python tools/ncu_target.py 888 3 > gpurun_out/r2_ncu.log 2>&1
tail -5 gpurun_out/r2_ncu.log
ls -la gpurun_out/
