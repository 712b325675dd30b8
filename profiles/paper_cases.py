"""Times the paper-size configurations (BASELINE.json configs[0..2]) through the drop-in wave_2D_solve call on one
GPU and the CPU oracle (factor-once) beside it.  These meshes are cache-resident (<= 44 k dofs): the numbers are
latency-, not bandwidth-bound.  Run: python profiles/paper_cases.py  (prints one JSON line per case)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from porthamiltonian_fem_b200 import wave_2D_solve
from porthamiltonian_fem_b200.mesh import rectangle_structured, generate_mesh
from oracle.wave_2D_oracle import wave_2D_solve_oracle

K, RHO = 3.0, 2.0
CASES = [
    ('config1 wave SE nx80-equivalent 136x34x2', rectangle_structured(136, 34), 1.0, 0.25, 1.5, 3000, dict(timeIntScheme='SE')),
    ('config2 IC SV nx120-equivalent 208x52x2', rectangle_structured(208, 52), 1.0, 0.25, 4.5, 3000, dict(timeIntScheme='SV', interConnection='IC')),
    ('config2 IC SM nx120-equivalent 208x52x2', rectangle_structured(208, 52), 1.0, 0.25, 4.5, 3000, dict(timeIntScheme='SM', interConnection='IC')),
    ('config3 passivity SM nx40-equivalent', generate_mesh('R', 40, 1.0, 0.25), 1.0, 0.25, 8.0, 8000,
     dict(timeIntScheme='SM', interConnection='IC', controlType='passivity', input_stop_t=0.0, control_start_t=1.0)),
    ('config3 casimir SM xLength=4 nx100-equivalent', generate_mesh('R', 100, 4.0, 0.25), 4.0, 0.25, 5.0, 5000,
     dict(timeIntScheme='SM', controlType='casimir')),
]
cpu_steps = int(os.environ.get('CPU_STEPS', '300'))
for name, mesh, xL, yL, tF, steps, kw in CASES:
    t0 = time.time()
    H, nc, st = wave_2D_solve(tF, steps, '/tmp', 0, xL, yL, K_wave=K, rho=RHO, mesh=mesh, return_state=True, verbose=False, **kw)
    wall = time.time() - t0
    stats = st['stats']
    ndof = len(st['p']) + len(st['q']) + (3 if kw.get('interConnection') else 0)
    timing = {}
    Ho, _ = wave_2D_solve_oracle(tF * cpu_steps / steps, cpu_steps, mesh, xL, yL, K_wave=K, rho=RHO, timing=timing, **kw)
    err = np.abs(H[:cpu_steps, 1:] - Ho[:, 1:]).max() / np.abs(Ho[:, 1:]).max()
    print(json.dumps({'case': name, 'cells': nc, 'dofs': ndof, 'steps': steps, 'gpu_device_ms': stats['device_ms'],
                      'gpu_steps_per_s': steps / (stats['device_ms'] * 1e-3), 'gpu_dof_updates_per_s': ndof * steps / (stats['device_ms'] * 1e-3),
                      'gpu_wall_s_incl_setup': wall, 'launches': stats['kernel_launches'],
                      'iters_p': stats['iters_p'] / max(1, stats['solves_p']), 'iters_q': stats['iters_q'] / max(1, stats['solves_q']),
                      'iters_block': stats['iters_block'] / max(1, stats['solves_block']),
                      'cpu_oracle_steps_per_s': cpu_steps / timing['loop_seconds'], 'H_trace_rel_diff_first_%d_steps' % cpu_steps: err,
                      'max_abs_E': float(np.abs(H[:, 2]).max()), 'H_end': float(H[-1, 1])}))
