#!/bin/bash
# round 2, GPU session T (1 GPU): sweep solve with two rows in flight per warp (libphwave_rows2_minb{4,3}.so) against the shipped kernels
# (libphwave_frozen.so = the in-tree build), both row mappings; then, with the shipped kernels: sweep parity tests, the whole GPU suite as
# the driver runs it, the default bench line
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for lib in frozen rows2_minb4 rows2_minb3; do
  for map in 0 1; do
    export PHW_LIB="$PWD/porthamiltonian_fem_b200/libphwave_$lib.so"
    ( time PHW_SWEEP_MAP=$map timeout 300 python bench.py --workload sweep --steps 100 --warmup 10 ) > gpurun_out/r2t_bench_sweep_${lib}_map$map.json 2> gpurun_out/r2t_bench_sweep_${lib}_map$map.err
    grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.]*' gpurun_out/r2t_bench_sweep_${lib}_map$map.json | tr '\n' ' '; echo " $lib map=$map"
  done
done
export PHW_LIB="$PWD/porthamiltonian_fem_b200/libphwave_rows2_minb3.so"
( time timeout 300 python -m pytest tests/test_gpu_loop.py -q -m gpu -p no:cacheprovider -k "sweep" ) > gpurun_out/r2t_sweep_tests_rows2.log 2>&1
echo "tests exit $?" >> gpurun_out/r2t_sweep_tests_rows2.log
tail -3 gpurun_out/r2t_sweep_tests_rows2.log | cut -c1-200
export PHW_LIB="$PWD/porthamiltonian_fem_b200/libphwave_frozen.so"
( time timeout 300 python bench.py --workload sweep --cases 4096 --steps 50 --warmup 10 ) > gpurun_out/r2t_bench_sweep_4096.json 2> gpurun_out/r2t_bench_sweep_4096.err
( time timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider ) > gpurun_out/r2t_gpu_suite.log 2>&1
echo "suite exit $?" >> gpurun_out/r2t_gpu_suite.log
tail -6 gpurun_out/r2t_gpu_suite.log | cut -c1-250
( time timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline ) > gpurun_out/r2t_bench_default.json 2> gpurun_out/r2t_bench_default.err
grep -o '"ms_per_step": [0-9.]*\|"frac": [0-9.]*' gpurun_out/r2t_bench_default.json | head -3
du -sh gpurun_out
