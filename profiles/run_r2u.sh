#!/bin/bash
# round 2, GPU session U (4 GPUs): partitioned parity on two and four ranks with the shipped binary (incl. the implicit scheme SM with the
# lumped model on four ranks), then the 4-GPU bench line (weak + strong) and the 4-GPU sweep
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi -L > gpurun_out/r2u_gpus.txt
( time timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -p no:cacheprovider -rs ) > gpurun_out/r2u_multi_parity_4gpus.log 2>&1
echo "tests exit $?" >> gpurun_out/r2u_multi_parity_4gpus.log
tail -8 gpurun_out/r2u_multi_parity_4gpus.log | cut -c1-250
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu-baseline ) > gpurun_out/r2u_bench_4gpu.log 2>&1
grep "^{" gpurun_out/r2u_bench_4gpu.log > gpurun_out/r2u_bench_4gpu.json
grep -o '"ms_per_step": [0-9.]*\|"efficiency": [0-9.]*\|"value": [0-9.e+]*' gpurun_out/r2u_bench_4gpu.json | head -6
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 4 --workload sweep --steps 100 --warmup 10 ) > gpurun_out/r2u_bench_sweep_4gpu.log 2>&1
grep "^{" gpurun_out/r2u_bench_sweep_4gpu.log > gpurun_out/r2u_bench_sweep_4gpu.json
grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.]*' gpurun_out/r2u_bench_sweep_4gpu.json | head -2
du -sh gpurun_out
