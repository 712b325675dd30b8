#!/bin/bash
# round 2, GPU session V (1 GPU, shipped binary): __graft_entry__.smoke(), the ncu launch list of the default bench command
# (--metrics gpu__time_duration.sum --clock-control none; shares of the step), the higher-order paper cases with the final default
# path selection, and one implicit higher-order case timed
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2v_smoke.log 2>&1
tail -4 gpurun_out/r2v_smoke.log
( time timeout 400 python bench.py --steps 2 --warmup 1 --no-cpu-baseline ) > gpurun_out/r2v_bench_steps2.log 2>&1
( time timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2v_launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline ) > gpurun_out/r2v_ncu_launches.log 2>&1
python profiles/summarize.py launches gpurun_out/r2v_launches_bench_steps2.csv > gpurun_out/r2v_launches_summary.txt 2>&1
tail -25 gpurun_out/r2v_launches_summary.txt | cut -c1-200
( time CPU_STEPS=20 timeout 400 python profiles/ho_cases.py ) > gpurun_out/r2v_ho_default.jsonl 2> gpurun_out/r2v_ho_default.err
python - <<'PY'
import json
for ln in open('gpurun_out/r2v_ho_default.jsonl'):
    if ln.startswith('{'):
        d = json.loads(ln)
        print('  %-46s %8d dofs %8.0f steps/s mf=%s its %.1f %.1f diff %.1e' % (d['case'], d['dofs'], d['gpu_steps_per_s'], d.get('matrix_free'), d['iters_p'], d['iters_q'], [v for k, v in d.items() if k.startswith('H_array_rel_diff')][0]))
PY
python - <<'PY' > gpurun_out/r2v_ho_implicit.jsonl 2>&1
import json, time, sys, os
sys.path.insert(0, os.getcwd())
from porthamiltonian_fem_b200 import wave_2D_solve
from porthamiltonian_fem_b200.mesh import rectangle_structured
for scheme, bo, (Nx, Ny), kw in [('SM', (2, 2), (136, 34), {}), ('SM', (3, 3), (136, 34), {}), ('SM', (2, 2), (136, 34), dict(interConnection='IC'))]:
    mesh = rectangle_structured(Nx, Ny)
    t0 = time.time()
    H, nc, st = wave_2D_solve(0.5, 1000, '/tmp', 0, 1.0, 0.25, K_wave=3.0, rho=2.0, mesh=mesh, timeIntScheme=scheme, basis_order=bo, return_state=True, verbose=False, **kw)
    s = st['stats']
    print(json.dumps({'case': 'R %s weak %dx%dx2 P%d/RT%d %s' % (scheme, Nx, Ny, bo[0], bo[1], kw.get('interConnection', 'wave')), 'dofs': len(st['p']) + len(st['q']), 'steps': 1000,
                      'gpu_steps_per_s': 1000 / (s['device_ms'] * 1e-3), 'iters_block': s['iters_block'] / max(1, s['solves_block']), 'max_rel_residual': s['max_rel_residual'],
                      'wall_s_incl_setup': time.time() - t0, 'H_end': float(H[-1, 1])}), flush=True)
PY
cat gpurun_out/r2v_ho_implicit.jsonl | cut -c1-300
du -sh gpurun_out
