#!/bin/bash
# round 2, GPU session F (8 GPUs): the default bench under torchrun at 8 ranks (weak headline + strong scaling of the ONE 50 M-dof
# mesh in config.strong, pipelined kernels with the fused peer exchange), then the batched sweep with one shard of 1024 cases per GPU
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r2f_gpus.txt
( time timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline ) > gpurun_out/r2f_bench_8gpu.log 2>&1
echo "bench exit $?" >> gpurun_out/r2f_bench_8gpu.log
tail -c 3500 gpurun_out/r2f_bench_8gpu.log
( time PHW_NO_LOOP=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --workload sweep --gpus 8 --steps 200 --warmup 10 ) > gpurun_out/r2f_sweep_8gpu.log 2>&1
echo "sweep exit $?" >> gpurun_out/r2f_sweep_8gpu.log
tail -c 900 gpurun_out/r2f_sweep_8gpu.log
