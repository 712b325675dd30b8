#!/bin/bash
# one short GPU session: parity of the new kernel variants first, then A/B bench lines (each step bounded by `timeout`)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/r1c_gpu.txt 2>&1
( time timeout 200 python -m pytest tests/test_gpu_pipe.py -x -q -p no:cacheprovider ) > gpurun_out/r1c_pipe_tests.log 2>&1
echo "pipe tests exit $?" >> gpurun_out/r1c_pipe_tests.log
tail -3 gpurun_out/r1c_pipe_tests.log
( time timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --pipe pq ) > gpurun_out/r1c_bench_pipe_pq.log 2>&1
tail -c 600 gpurun_out/r1c_bench_pipe_pq.log
( time timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --pipe p ) > gpurun_out/r1c_bench_pipe_p.log 2>&1
tail -c 300 gpurun_out/r1c_bench_pipe_p.log
( time timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/r1c_bench_default.log 2>&1
tail -c 300 gpurun_out/r1c_bench_default.log
