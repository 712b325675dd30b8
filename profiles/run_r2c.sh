#!/bin/bash
# round 2, GPU session C (exchange protocol v2) (2 GPUs): partitioned-run parity for lean AND pipelined kernels (10 cases), then the bench under torchrun
# (weak headline + strong scaling of the one 50 M-dof mesh in config.strong)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r2c_gpus.txt
( time timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider -rs ) > gpurun_out/r2c_multi_parity.log 2>&1
echo "multi parity exit $?" >> gpurun_out/r2c_multi_parity.log
tail -15 gpurun_out/r2c_multi_parity.log
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline ) > gpurun_out/r2c_bench_2gpu.log 2>&1
echo "bench exit $?" >> gpurun_out/r2c_bench_2gpu.log
tail -c 3000 gpurun_out/r2c_bench_2gpu.log
