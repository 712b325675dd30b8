#!/bin/bash
# round 2, GPU session K (8 GPUs): 4-rank partitioned parity (interior slabs with two neighbours), the default bench at 8 ranks (weak headline
# + strong scaling of the ONE 50 M-dof mesh), the same at 4 ranks, and the batched sweep with one shard of 1024 cases per GPU
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r2k_gpus.txt
( time PHW_TEST_PIPE_MULTI=1 timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider -rs -k four ) > gpurun_out/r2k_multi_parity_4ranks.log 2>&1
echo "multi parity exit $?" >> gpurun_out/r2k_multi_parity_4ranks.log
tail -6 gpurun_out/r2k_multi_parity_4ranks.log
( time timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline ) > gpurun_out/r2k_bench_8gpu.log 2>&1
echo "bench exit $?" >> gpurun_out/r2k_bench_8gpu.log
tail -c 3500 gpurun_out/r2k_bench_8gpu.log
( time timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline ) > gpurun_out/r2k_bench_1gpu.log 2>&1
tail -c 600 gpurun_out/r2k_bench_1gpu.log
( time PHW_NO_LOOP=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --workload sweep --gpus 8 --steps 200 --warmup 10 --patch-cells 1024 --sweep-pipe --pipe pq ) > gpurun_out/r2k_sweep_8gpu.log 2>&1
echo "sweep exit $?" >> gpurun_out/r2k_sweep_8gpu.log
tail -c 900 gpurun_out/r2k_sweep_8gpu.log
