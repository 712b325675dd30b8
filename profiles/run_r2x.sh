#!/bin/bash
# round 2, GPU session X (1 GPU, in-tree build): all sweep kernels templated on the cases per lane (right-hand sides and energies too) -
# 1 / 2 / 4 cases per lane on the 1024- and 4096-case sweeps, sweep parity tests at 2 (default) and 4 cases per lane, the whole GPU suite
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for cpl in 2 4 1; do
  ( time PHW_SWEEP_CPL=$cpl timeout 300 python bench.py --workload sweep --steps 100 --warmup 10 ) > gpurun_out/r2x_bench_sweep_cpl$cpl.json 2> gpurun_out/r2x_bench_sweep_cpl$cpl.err
  grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.]*' gpurun_out/r2x_bench_sweep_cpl$cpl.json | tr '\n' ' '; echo " cpl=$cpl"
done
for cpl in 2 4; do
  ( time PHW_SWEEP_CPL=$cpl timeout 300 python bench.py --workload sweep --cases 4096 --steps 50 --warmup 10 ) > gpurun_out/r2x_bench_sweep_4096_cpl$cpl.json 2> gpurun_out/r2x_bench_sweep_4096_cpl$cpl.err
  grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.]*' gpurun_out/r2x_bench_sweep_4096_cpl$cpl.json | tr '\n' ' '; echo " 4096 cases cpl=$cpl"
done
( time PHW_SWEEP_CPL=4 timeout 300 python -m pytest tests/test_gpu_loop.py -q -m gpu -p no:cacheprovider -k "sweep" ) > gpurun_out/r2x_sweep_tests_cpl4.log 2>&1
echo "tests exit $?" >> gpurun_out/r2x_sweep_tests_cpl4.log
tail -3 gpurun_out/r2x_sweep_tests_cpl4.log | cut -c1-200
( time timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider ) > gpurun_out/r2x_gpu_suite.log 2>&1
echo "suite exit $?" >> gpurun_out/r2x_gpu_suite.log
tail -6 gpurun_out/r2x_gpu_suite.log | cut -c1-250
( time timeout 300 ncu --set full --clock-control none -k 'regex:solve_coop_csr_b_k' --launch-skip 22 -c 2 -f -o /tmp/r2x_sweep python bench.py --workload sweep --steps 4 --warmup 10 ) > gpurun_out/r2x_ncu_sweep.log 2>&1
python profiles/ncu_extract.py /tmp/r2x_sweep.ncu-rep gpurun_out/r2x_ncu_full_sweep_solve.txt > /dev/null 2>> gpurun_out/r2x_ncu_sweep.log
ncu -i /tmp/r2x_sweep.ncu-rep --page details 2>/dev/null | grep -v "^ *$" | head -230 > gpurun_out/r2x_ncu_sweep_solve_details.txt
cat gpurun_out/r2x_ncu_full_sweep_solve.txt | cut -c1-220
du -sh gpurun_out
