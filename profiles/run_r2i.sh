#!/bin/bash
# round 2, GPU session I (1 GPU): resident solves v3 (Jacobi-prescaled rows: no division, no diagonal in the iteration; config 2 fits again), tests, paper cases, sweep
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 600 python -m pytest tests/test_gpu_loop.py tests/test_gpu_field_output.py -m gpu -x -q -p no:cacheprovider ) > gpurun_out/r2i_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2i_tests.log
tail -25 gpurun_out/r2i_tests.log
( time CPU_STEPS=300 timeout 400 python profiles/paper_cases.py ) > gpurun_out/r2i_paper_res.jsonl 2> gpurun_out/r2i_paper_res.err
python - <<'PY'
import json
for f in ('gpurun_out/r2i_paper_res.jsonl',):
    print(f)
    for ln in open(f):
        if ln.startswith('{'):
            d = json.loads(ln)
            print('  %-50s %8.0f steps/s  %7.1f us/step  loop=%s res=%s launches=%d its %.1f %.1f %.1f cpu %6.0f steps/s  diff %.1e' % (d['case'], d['gpu_steps_per_s'], d['gpu_us_per_step'], d['whole_loop_kernel'], d.get('resident_solves'), d['launches'], d['iters_p'], d['iters_q'], d['iters_block'], d.get('cpu_oracle_steps_per_s', 0), d.get('H_trace_rel_diff_first_300_steps', -1)))
PY
tail -3 gpurun_out/r2i_paper_res.err
for v in "PHW_LOOP_CL=4" "PHW_NO_LOOP=1"; do
  tag=$(echo "$v" | tr ' =' '__'); [ -z "$tag" ] && tag=default
  ( time env $v timeout 300 python bench.py --workload sweep --steps 200 --warmup 10 ) > gpurun_out/r2i_sweep_$tag.log 2>&1
  echo "sweep [$v]: $(grep -o '"value": [0-9.e+]*' gpurun_out/r2i_sweep_$tag.log | head -1) $(grep -o '"ms_per_step": [0-9.e+]*' gpurun_out/r2i_sweep_$tag.log | head -1) $(grep -o '"iters_per_solve_p": [0-9.]*, "iters_per_solve_q": [0-9.]*' gpurun_out/r2i_sweep_$tag.log) $(grep -o '"gpu_launches": [0-9]*' gpurun_out/r2i_sweep_$tag.log) $(grep -c Error gpurun_out/r2i_sweep_$tag.log) errors"
done
