#!/bin/bash
# round 2, GPU session E (1 GPU): the single-launch whole-loop kernel - parity suite, the whole GPU suite, the paper cases and the batched
# sweep with it and (PHW_NO_LOOP=1) with the multi-launch path of round 1, then the default bench line
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 600 python -m pytest tests/test_gpu_loop.py -m gpu -x -q -p no:cacheprovider ) > gpurun_out/r2e_loop_tests.log 2>&1
echo "loop tests exit $?" >> gpurun_out/r2e_loop_tests.log
tail -25 gpurun_out/r2e_loop_tests.log
( time timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_loop.py ) > gpurun_out/r2e_gpu_suite.log 2>&1
echo "suite exit $?" >> gpurun_out/r2e_gpu_suite.log
tail -6 gpurun_out/r2e_gpu_suite.log
( time timeout 400 python profiles/paper_cases.py ) > gpurun_out/r2e_paper_loop.jsonl 2> gpurun_out/r2e_paper_loop.err
( time PHW_NO_LOOP=1 timeout 400 python profiles/paper_cases.py ) > gpurun_out/r2e_paper_multilaunch.jsonl 2> gpurun_out/r2e_paper_multilaunch.err
python - <<'PY'
import json
for f in ('gpurun_out/r2e_paper_loop.jsonl', 'gpurun_out/r2e_paper_multilaunch.jsonl'):
    print(f)
    for ln in open(f):
        if ln.startswith('{'):
            d = json.loads(ln)
            print('  %-50s %8.0f steps/s  %7.1f us/step  loop=%s launches=%d  cpu %6.0f steps/s  diff %.1e' % (d['case'], d['gpu_steps_per_s'], d['gpu_us_per_step'], d['whole_loop_kernel'], d['launches'], d.get('cpu_oracle_steps_per_s', 0), d.get('H_trace_rel_diff_first_300_steps', -1)))
PY
tail -3 gpurun_out/r2e_paper_loop.err
( time timeout 400 python bench.py --workload sweep --steps 200 --warmup 10 ) > gpurun_out/r2e_sweep_loop.log 2>&1
tail -c 900 gpurun_out/r2e_sweep_loop.log
( time PHW_NO_LOOP=1 timeout 400 python bench.py --workload sweep --steps 200 --warmup 10 ) > gpurun_out/r2e_sweep_multilaunch.log 2>&1
tail -c 900 gpurun_out/r2e_sweep_multilaunch.log
( time timeout 300 python bench.py --no-cpu-baseline ) > gpurun_out/r2e_bench_default.log 2>&1
tail -c 500 gpurun_out/r2e_bench_default.log
