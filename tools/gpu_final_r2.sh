# round-2 evidence: ncu captures of the final kernels, DRAM traffic of one profiled batch, launch list of the bench command, bench lines
mkdir -p gpurun_out
export ESLAM_SYM_GRAPH=0
# second batch of tools/ncu_target.py 888 2 (888 objects = 2 waves of 444 CTAs): extract_kernel + the first rounds of the symmetry pipeline
# (step0 eval0 step1 eval1 step2 eval2 step3 eval3 step4); a batch has 1 + 57 + 56 launches matching the filter
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'extract_kernel|sym_eval|sym_step' -s 114 -c 10 -o gpurun_out/rf2_prof -f python tools/ncu_target.py 888 2 > gpurun_out/rf2_ncu.log 2>&1
tail -2 gpurun_out/rf2_ncu.log
# time + DRAM bytes of every launch of one profiled batch (caches NOT flushed between launches: the seeds' state is meant to stay in L2)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -k regex:'extract_kernel|sym_|symmetry' -s 116 -c 116 --csv --log-file gpurun_out/rf2_traffic_launches.csv python tools/ncu_target.py 888 2 > gpurun_out/rf2_ncu_traffic.log 2>&1
python tools/ncu_kernel_table.py gpurun_out/rf2_traffic_launches.csv 888 | tee gpurun_out/rf2_traffic_table.txt
unset ESLAM_SYM_GRAPH
# launch list of the bench command (graphs on: kernel nodes are listed like launches)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/rf2_launches_bench.csv python bench.py --steps 2 --warmup 3 --frames 512 --cpu-sample 8 --no-stages > gpurun_out/rf2_bench_under_ncu.log 2>&1
tail -1 gpurun_out/rf2_bench_under_ncu.log | cut -c1-120
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/rf2_bench_n1.json 2> gpurun_out/rf2_bench.err
cut -c1-200 gpurun_out/rf2_bench_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/rf2_bench_reference.json 2>> gpurun_out/rf2_bench.err
cut -c1-300 gpurun_out/rf2_bench_reference.json
python tools/extract_phases.py > gpurun_out/rf2_phase_clocks.txt 2>&1; ESLAM_SYM_IMPL=2 python tools/extract_phases.py 2>&1 | tail -4 >> gpurun_out/rf2_phase_clocks.txt; cat gpurun_out/rf2_phase_clocks.txt
bash tools/gpu_validate.sh
