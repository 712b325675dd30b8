#!/bin/bash
# final GPU session of the round: whole GPU suite, default bench line, ncu launch list of the same command, ncu --set full of the solves
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 150 python -m pytest tests -m gpu -x -q -p no:cacheprovider ) > gpurun_out/r1h_gpu_suite.log 2>&1
echo "suite exit $?" >> gpurun_out/r1h_gpu_suite.log
tail -4 gpurun_out/r1h_gpu_suite.log
( time timeout 120 python bench.py ) > gpurun_out/r1h_bench_default.log 2>&1
tail -c 400 gpurun_out/r1h_bench_default.log
( time timeout 70 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1h_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline ) > gpurun_out/r1h_ncu_list.log 2>&1
tail -2 gpurun_out/r1h_ncu_list.log
( time timeout 45 ncu --set full --clock-control none --import-source on -k regex:solve_.*pipe_k -c 2 -o gpurun_out/r1h_pipe_solves python bench.py --steps 1 --warmup 0 --no-cpu-baseline ) > gpurun_out/r1h_ncu_full.log 2>&1
tail -2 gpurun_out/r1h_ncu_full.log
