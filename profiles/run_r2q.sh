#!/bin/bash
# round 2, GPU session Q (1 GPU): the batched sweep with the case index fastest (csrc/kernels_sweep.cuh) - parity tests, the sweep bench
# line with the new layout and with the replicated-mesh layout beside it, ncu --set full of its solve kernel - then the whole GPU suite
# exactly as the driver runs it
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
[ -f porthamiltonian_fem_b200/libphwave_frozen.so ] && export PHW_LIB="$PWD/porthamiltonian_fem_b200/libphwave_frozen.so"
( time timeout 600 python -m pytest tests/test_gpu_loop.py -q -m gpu -p no:cacheprovider -k "sweep" ) > gpurun_out/r2q_sweep_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2q_sweep_tests.log
tail -30 gpurun_out/r2q_sweep_tests.log | cut -c1-250
( time timeout 300 python bench.py --workload sweep --steps 100 --warmup 10 ) > gpurun_out/r2q_bench_sweep.json 2> gpurun_out/r2q_bench_sweep.err
( time PHW_SWEEP_GROUPED=1 timeout 300 python bench.py --workload sweep --steps 100 --warmup 10 ) > gpurun_out/r2q_bench_sweep_grouped.json 2> gpurun_out/r2q_bench_sweep_grouped.err
( time timeout 300 python bench.py --workload sweep --cases 4096 --steps 50 --warmup 10 ) > gpurun_out/r2q_bench_sweep_4096.json 2> gpurun_out/r2q_bench_sweep_4096.err
grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.]*\|"iters_per_solve_[pq]": [0-9.]*' gpurun_out/r2q_bench_sweep.json gpurun_out/r2q_bench_sweep_grouped.json gpurun_out/r2q_bench_sweep_4096.json
tail -3 gpurun_out/r2q_bench_sweep.err
( time timeout 300 ncu --set full --clock-control none -k 'regex:solve_coop_csr_b_k|csr_store_b_k|csr_dot_b_k' -c 6 -f -o /tmp/r2q_sweep_full python bench.py --workload sweep --steps 2 --warmup 1 ) > gpurun_out/r2q_ncu_sweep.log 2>&1
python profiles/ncu_extract.py /tmp/r2q_sweep_full.ncu-rep gpurun_out/r2q_ncu_full_sweep_kernels.txt > /dev/null 2>> gpurun_out/r2q_ncu_sweep.log
ncu -i /tmp/r2q_sweep_full.ncu-rep --page details 2>/dev/null | grep -v "^ *$" | awk '/solve_coop_csr_b_k/{c++} c>=1' | head -260 > gpurun_out/r2q_ncu_sweep_solve_details.txt
cat gpurun_out/r2q_ncu_full_sweep_kernels.txt | cut -c1-220
( time timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider ) > gpurun_out/r2q_gpu_suite.log 2>&1
echo "suite exit $?" >> gpurun_out/r2q_gpu_suite.log
tail -8 gpurun_out/r2q_gpu_suite.log | cut -c1-250
du -sh gpurun_out
