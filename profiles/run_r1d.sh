#!/bin/bash
# second short GPU session: parity of all pipelined variants, A/B bench lines, then the whole GPU suite with default flags
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 240 python -m pytest tests/test_gpu_pipe.py -x -q -p no:cacheprovider ) > gpurun_out/r1d_pipe_tests.log 2>&1
echo "pipe tests exit $?" >> gpurun_out/r1d_pipe_tests.log
tail -4 gpurun_out/r1d_pipe_tests.log
( time timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --pipe pq ) > gpurun_out/r1d_bench_pipe_pq_diag.log 2>&1
tail -c 300 gpurun_out/r1d_bench_pipe_pq_diag.log
( time timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --pipe pqer ) > gpurun_out/r1d_bench_pipe_pqer.log 2>&1
tail -c 300 gpurun_out/r1d_bench_pipe_pqer.log
( time timeout 400 python -m pytest tests -m gpu -x -q -p no:cacheprovider --deselect tests/test_gpu_pipe.py ) > gpurun_out/r1d_gpu_suite.log 2>&1
echo "suite exit $?" >> gpurun_out/r1d_gpu_suite.log
tail -4 gpurun_out/r1d_gpu_suite.log
