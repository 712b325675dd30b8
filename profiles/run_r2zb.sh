#!/bin/bash
# round 2, GPU session ZB (8 GPUs, final tree): the 8-GPU bench line (weak + strong) and the 8-GPU sweep with the case index fastest
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi -L | wc -l > gpurun_out/r2zb_gpus.txt
( time timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline ) > gpurun_out/r2zb_bench_8gpu.log 2>&1
grep "^{" gpurun_out/r2zb_bench_8gpu.log > gpurun_out/r2zb_bench_8gpu.json
grep -o '"ms_per_step": [0-9.]*\|"efficiency": [0-9.]*\|"value": [0-9.e+]*' gpurun_out/r2zb_bench_8gpu.json | head -6
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --workload sweep --steps 100 --warmup 10 ) > gpurun_out/r2zb_bench_sweep_8gpu.log 2>&1
grep "^{" gpurun_out/r2zb_bench_sweep_8gpu.log > gpurun_out/r2zb_bench_sweep_8gpu.json
grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.]*' gpurun_out/r2zb_bench_sweep_8gpu.json | head -2
