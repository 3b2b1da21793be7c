mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r9_bench_n1.json 2> gpurun_out/r9_bench.err
cat gpurun_out/r9_bench_n1.json; tail -2 gpurun_out/r9_bench.err
timeout 600 python bench.py --steps 5 --warmup 3 --frames 256 --workload multi_box --cpu-sample 64 > gpurun_out/r9_bench_multibox.json 2>> gpurun_out/r9_bench.err
cat gpurun_out/r9_bench_multibox.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r9_bench_reference_arm.json 2>> gpurun_out/r9_bench.err
cat gpurun_out/r9_bench_reference_arm.json
timeout 600 python tools/bench_stages.py > gpurun_out/r9_stages.json 2>> gpurun_out/r9_bench.err
cat gpurun_out/r9_stages.json
