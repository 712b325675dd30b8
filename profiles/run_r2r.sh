#!/bin/bash
# round 2, GPU session R (1 GPU): case-index-fastest sweep, second version (cyclic row tiles: the CTAs of a case chunk sweep one window of
# consecutive rows, so neighbour rows hit in L2) at three occupancy targets of the persistent solve (libphwave_minb{5,6,8}.so: 48 / 40 / 32
# registers), sweep parity tests with the chosen mapping, ncu --set full of representative solve launches
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for mb in 5 6 8; do
  export PHW_LIB="$PWD/porthamiltonian_fem_b200/libphwave_minb$mb.so"
  ( time timeout 300 python bench.py --workload sweep --steps 100 --warmup 10 ) > gpurun_out/r2r_bench_sweep_minb$mb.json 2> gpurun_out/r2r_bench_sweep_minb$mb.err
  grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.]*' gpurun_out/r2r_bench_sweep_minb$mb.json | tr '\n' ' '; echo " minb=$mb"
done
export PHW_LIB="$PWD/porthamiltonian_fem_b200/libphwave_minb5.so"
( time timeout 300 python bench.py --workload sweep --cases 4096 --steps 50 --warmup 10 ) > gpurun_out/r2r_bench_sweep_4096_minb5.json 2> gpurun_out/r2r_bench_sweep_4096_minb5.err
grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.]*' gpurun_out/r2r_bench_sweep_4096_minb5.json | tr '\n' ' '; echo " 4096 cases minb=5"
( time timeout 600 python -m pytest tests/test_gpu_loop.py -q -m gpu -p no:cacheprovider -k "sweep" ) > gpurun_out/r2r_sweep_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2r_sweep_tests.log
tail -5 gpurun_out/r2r_sweep_tests.log | cut -c1-250
for mb in 5 8; do
  export PHW_LIB="$PWD/porthamiltonian_fem_b200/libphwave_minb$mb.so"
  ( time timeout 300 ncu --set full --clock-control none -k 'regex:solve_coop_csr_b_k' --launch-skip 22 -c 2 -f -o /tmp/r2r_sweep_minb$mb python bench.py --workload sweep --steps 4 --warmup 10 ) > gpurun_out/r2r_ncu_sweep_minb$mb.log 2>&1
  python profiles/ncu_extract.py /tmp/r2r_sweep_minb$mb.ncu-rep gpurun_out/r2r_ncu_full_sweep_solve_minb$mb.txt > /dev/null 2>> gpurun_out/r2r_ncu_sweep_minb$mb.log
  ncu -i /tmp/r2r_sweep_minb$mb.ncu-rep --page details 2>/dev/null | grep -v "^ *$" | head -230 > gpurun_out/r2r_ncu_sweep_solve_details_minb$mb.txt
  cat gpurun_out/r2r_ncu_full_sweep_solve_minb$mb.txt | cut -c1-220
done
du -sh gpurun_out
