#!/bin/bash
# round 2, GPU session ZE (1 GPU, FINAL build of the final tree): the whole GPU suite as the driver runs it, smoke(), ncu --set full of the
# shipped sweep solve (four cases per lane), 3 000 + 3 000 more full-size steps of the default bench (stress of the pipelined path)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider ) > gpurun_out/r2ze_gpu_suite.log 2>&1
echo "suite exit $?" >> gpurun_out/r2ze_gpu_suite.log
tail -6 gpurun_out/r2ze_gpu_suite.log | cut -c1-250
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2ze_smoke.log 2>&1; grep "smoke ok" gpurun_out/r2ze_smoke.log
( time timeout 300 ncu --set full --clock-control none -k 'regex:solve_coop_csr_b_k' --launch-skip 22 -c 2 -f -o /tmp/r2ze_sweep python bench.py --workload sweep --steps 4 --warmup 10 ) > gpurun_out/r2ze_ncu_sweep.log 2>&1
python profiles/ncu_extract.py /tmp/r2ze_sweep.ncu-rep gpurun_out/r2ze_ncu_full_sweep_solve_four_cases_per_lane.txt > /dev/null 2>> gpurun_out/r2ze_ncu_sweep.log
ncu -i /tmp/r2ze_sweep.ncu-rep --page details 2>/dev/null | grep -v "^ *$" | head -230 > gpurun_out/r2ze_ncu_sweep_solve_details.txt
cat gpurun_out/r2ze_ncu_full_sweep_solve_four_cases_per_lane.txt | cut -c1-220
( time timeout 600 python bench.py --steps 3000 --warmup 3 --no-cpu-baseline ) > gpurun_out/r2ze_stress_6000_steps.json 2> gpurun_out/r2ze_stress_6000_steps.err
echo "stress exit $?" >> gpurun_out/r2ze_stress_6000_steps.err
grep -o '"ms_per_step": [0-9.]*\|"max_rel_residual": [0-9.e-]*\|"p_nodal_rel_l2": [0-9.e-]*' gpurun_out/r2ze_stress_6000_steps.json | head -4
du -sh gpurun_out
