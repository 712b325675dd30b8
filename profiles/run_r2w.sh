#!/bin/bash
# round 2, GPU session W (1 GPU): sweep solve with two consecutive cases per lane (16-byte accesses; in-tree build: compiled for 4 CTAs per
# SM, libphwave_cpl2_minb3.so: 3) against one case per lane (PHW_SWEEP_CPL=1), both row mappings; sweep parity tests with the new
# default (B = 40: two per lane, B = 33: one); the whole GPU suite as the driver runs it
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
cp porthamiltonian_fem_b200/libphwave.so /tmp/libphwave_intree.so
for cfg in "intree 2" "cpl2_minb3 2" "intree 1"; do
  set -- $cfg
  [ $1 = intree ] && export PHW_LIB=/tmp/libphwave_intree.so || export PHW_LIB="$PWD/porthamiltonian_fem_b200/libphwave_$1.so"
  for map in 0 1; do
    ( time PHW_SWEEP_CPL=$2 PHW_SWEEP_MAP=$map timeout 300 python bench.py --workload sweep --steps 100 --warmup 10 ) > gpurun_out/r2w_bench_sweep_$1_cpl$2_map$map.json 2> gpurun_out/r2w_bench_sweep_$1_cpl$2_map$map.err
    grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.]*' gpurun_out/r2w_bench_sweep_$1_cpl$2_map$map.json | tr '\n' ' '; echo " $1 cpl=$2 map=$map"
  done
done
unset PHW_LIB
( time timeout 300 python bench.py --workload sweep --cases 4096 --steps 50 --warmup 10 ) > gpurun_out/r2w_bench_sweep_4096.json 2> gpurun_out/r2w_bench_sweep_4096.err
grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.]*' gpurun_out/r2w_bench_sweep_4096.json | tr '\n' ' '; echo " 4096 cases"
( time timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider ) > gpurun_out/r2w_gpu_suite.log 2>&1
echo "suite exit $?" >> gpurun_out/r2w_gpu_suite.log
tail -6 gpurun_out/r2w_gpu_suite.log | cut -c1-250
du -sh gpurun_out
