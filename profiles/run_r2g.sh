#!/bin/bash
# round 2, GPU session G (1 GPU): the whole-loop kernel with the mass solves resident in shared memory (RES: DSMEM ghost pulls, one
# cluster barrier per Chebyshev iteration) - parity suite, paper cases, batched sweep (cluster sizes 8 / 4, 512 / 1024 threads),
# then 8 000 more full-size steps of the default bench (stress of the pipelined path: 12 000 + 8 000 = 20 000)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 600 python -m pytest tests/test_gpu_loop.py -m gpu -x -q -p no:cacheprovider ) > gpurun_out/r2g_loop_tests.log 2>&1
echo "loop tests exit $?" >> gpurun_out/r2g_loop_tests.log
tail -25 gpurun_out/r2g_loop_tests.log
( time CPU_STEPS=300 timeout 400 python profiles/paper_cases.py ) > gpurun_out/r2g_paper_res.jsonl 2> gpurun_out/r2g_paper_res.err
( time CPU_STEPS=0 PHW_RES_NT=1024 timeout 400 python profiles/paper_cases.py ) > gpurun_out/r2g_paper_res1024.jsonl 2> gpurun_out/r2g_paper_res1024.err
python - <<'PY'
import json
for f in ('gpurun_out/r2g_paper_res.jsonl', 'gpurun_out/r2g_paper_res1024.jsonl'):
    print(f)
    for ln in open(f):
        if ln.startswith('{'):
            d = json.loads(ln)
            print('  %-50s %8.0f steps/s  %7.1f us/step  loop=%s res=%s launches=%d its %.1f %.1f %.1f cpu %6.0f steps/s  diff %.1e' % (d['case'], d['gpu_steps_per_s'], d['gpu_us_per_step'], d['whole_loop_kernel'], d.get('resident_solves'), d['launches'], d['iters_p'], d['iters_q'], d['iters_block'], d.get('cpu_oracle_steps_per_s', 0), d.get('H_trace_rel_diff_first_300_steps', -1)))
PY
tail -3 gpurun_out/r2g_paper_res.err
for v in "" "PHW_LOOP_CL=4" "PHW_RES_NT=1024" "PHW_LOOP_CL=4 PHW_RES_NT=1024" "PHW_NO_RES=1"; do
  tag=$(echo "$v" | tr ' =' '__'); [ -z "$tag" ] && tag=default
  ( time env $v timeout 300 python bench.py --workload sweep --steps 200 --warmup 10 ) > gpurun_out/r2g_sweep_$tag.log 2>&1
  echo "sweep [$v]: $(grep -o '"value": [0-9.e+]*' gpurun_out/r2g_sweep_$tag.log | head -1) $(grep -o '"ms_per_step": [0-9.e+]*' gpurun_out/r2g_sweep_$tag.log | head -1) $(grep -o '"iters_per_solve_p": [0-9.]*, "iters_per_solve_q": [0-9.]*' gpurun_out/r2g_sweep_$tag.log) $(grep -o '"gpu_launches": [0-9]*' gpurun_out/r2g_sweep_$tag.log) $(grep -c Error gpurun_out/r2g_sweep_$tag.log) errors"
done
( time timeout 600 python bench.py --steps 4000 --no-cpu-baseline ) > gpurun_out/r2g_stress8000.log 2>&1
echo "stress exit $?" >> gpurun_out/r2g_stress8000.log
tail -c 700 gpurun_out/r2g_stress8000.log
