mkdir -p gpurun_out
echo "NCCL_DEBUG=${NCCL_DEBUG:-unset}"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --cpu-sample 16 > gpurun_out/r20_n2.out 2> gpurun_out/r20_n2.err
echo "stdout lines: $(grep -c . gpurun_out/r20_n2.out)"; cut -c1-80 gpurun_out/r20_n2.out
echo "NCCL banner in stderr: $(grep -c 'NCCL version' gpurun_out/r20_n2.err)"
