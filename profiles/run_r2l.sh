#!/bin/bash
# round 2, GPU session L (2 GPUs): eigenvalue + sweep tests on GPU 0, then the exchange timing diagnostics (PHW_XCH_DEBUG=64: globaltimer
# stamps inside the pipelined solves) on the weak slab and on a slab of the size of the 8-rank strong-scaling run
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 600 python -m pytest tests/test_gpu_eigen.py tests/test_gpu_loop.py -m gpu -x -q -p no:cacheprovider -k "eigen or spectrum or second_order or sweep" ) > gpurun_out/r2l_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2l_tests.log
tail -12 gpurun_out/r2l_tests.log
( time PHW_XCH_DEBUG=64 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-strong ) > gpurun_out/r2l_diag_weak.log 2>&1
grep "^rank" gpurun_out/r2l_diag_weak.log | cut -c1-900
( time PHW_XCH_DEBUG=64 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --nx 884 --steps 20 --warmup 5 --no-cpu-baseline --no-strong ) > gpurun_out/r2l_diag_small.log 2>&1
grep "^rank" gpurun_out/r2l_diag_small.log | cut -c1-900
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2l_diag_small.log | head -2
( time timeout 300 python bench.py --gpus 1 --nx 884 --steps 20 --warmup 5 --no-cpu-baseline ) > gpurun_out/r2l_small_1gpu.log 2>&1
grep -o '"ms_per_step": [0-9.]*\|"ms_per_iter_[pq]": [0-9.]*' gpurun_out/r2l_small_1gpu.log | head -4
