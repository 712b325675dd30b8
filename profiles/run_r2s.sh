#!/bin/bash
# round 2, GPU session S (1 GPU): case-index-fastest sweep, third version (row indices / values loaded once per warp and broadcast by
# shuffles, four gathers in flight per lane) - both row mappings (PHW_SWEEP_MAP=1 cyclic tiles / 0 contiguous blocks) at three occupancy
# targets of the persistent solve (libphwave_minb{4,5,6}.so: 64 / 48 / 40 registers), sweep parity tests, ncu of representative launches
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for mb in 4 5 6; do
  for map in 1 0; do
    export PHW_LIB="$PWD/porthamiltonian_fem_b200/libphwave_minb$mb.so"
    ( time PHW_SWEEP_MAP=$map timeout 300 python bench.py --workload sweep --steps 100 --warmup 10 ) > gpurun_out/r2s_bench_sweep_minb${mb}_map$map.json 2> gpurun_out/r2s_bench_sweep_minb${mb}_map$map.err
    grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.]*' gpurun_out/r2s_bench_sweep_minb${mb}_map$map.json | tr '\n' ' '; echo " minb=$mb map=$map"
  done
done
export PHW_LIB="$PWD/porthamiltonian_fem_b200/libphwave_minb5.so"
for map in 1 0; do
  ( time PHW_SWEEP_MAP=$map timeout 600 python -m pytest tests/test_gpu_loop.py -q -m gpu -p no:cacheprovider -k "sweep" ) > gpurun_out/r2s_sweep_tests_map$map.log 2>&1
  echo "tests exit $?" >> gpurun_out/r2s_sweep_tests_map$map.log
  tail -4 gpurun_out/r2s_sweep_tests_map$map.log | cut -c1-250
done
for map in 1 0; do
  ( time PHW_SWEEP_MAP=$map timeout 300 ncu --set full --clock-control none -k 'regex:solve_coop_csr_b_k' --launch-skip 22 -c 2 -f -o /tmp/r2s_sweep_map$map python bench.py --workload sweep --steps 4 --warmup 10 ) > gpurun_out/r2s_ncu_sweep_map$map.log 2>&1
  python profiles/ncu_extract.py /tmp/r2s_sweep_map$map.ncu-rep gpurun_out/r2s_ncu_full_sweep_solve_map$map.txt > /dev/null 2>> gpurun_out/r2s_ncu_sweep_map$map.log
  ncu -i /tmp/r2s_sweep_map$map.ncu-rep --page details 2>/dev/null | grep -v "^ *$" | head -230 > gpurun_out/r2s_ncu_sweep_solve_details_map$map.txt
  cat gpurun_out/r2s_ncu_full_sweep_solve_map$map.txt | cut -c1-220
done
du -sh gpurun_out
