#!/bin/bash
# round 2, GPU session A (1 GPU): state of the round-1 binary - whole GPU suite, the 17 experimental (two-sweep pipelined) cases,
# default bench line, the pipe2 A/B, a 6000-step stress of the default pipelined path, ncu --set full of every pipelined kernel
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 400 python -m pytest tests -m gpu -x -q -p no:cacheprovider ) > gpurun_out/r2a_gpu_suite.log 2>&1
echo "suite exit $?" >> gpurun_out/r2a_gpu_suite.log
tail -4 gpurun_out/r2a_gpu_suite.log
( time PHW_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_pipe.py -m gpu -q -p no:cacheprovider ) > gpurun_out/r2a_experimental.log 2>&1
echo "experimental exit $?" >> gpurun_out/r2a_experimental.log
tail -15 gpurun_out/r2a_experimental.log
( time timeout 200 python bench.py ) > gpurun_out/r2a_bench_default.log 2>&1
tail -c 600 gpurun_out/r2a_bench_default.log
for cfg in PQer pQer Pqer; do
  ( time timeout 150 python bench.py --pipe $cfg --steps 10 --no-cpu-baseline ) > gpurun_out/r2a_bench_$cfg.log 2>&1
  tail -c 300 gpurun_out/r2a_bench_$cfg.log
done
( time PHW_BENCH_NO_FALLBACK=1 timeout 400 python bench.py --steps 6000 --no-cpu-baseline ) > gpurun_out/r2a_stress6000.log 2>&1
echo "stress exit $?" >> gpurun_out/r2a_stress6000.log
tail -c 400 gpurun_out/r2a_stress6000.log
( time timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:pipe_k|axpy_k|predict_k|copy_from_tb_k' -c 20 -f -o gpurun_out/r2a_pipe_full python bench.py --steps 2 --warmup 0 --no-cpu-baseline ) > gpurun_out/r2a_ncu_full.log 2>&1
tail -3 gpurun_out/r2a_ncu_full.log
ls -la gpurun_out | tail -5
