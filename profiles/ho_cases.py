"""(PHW_HO_MF=1 / PHW_HO_CSR=1 force the matrix-free / the assembled-CSR path; default: matrix-free at order 3.)
Times the reference's default caseArray (main_wave_2D.py:105-114: R, SV / SE, weak, nx80, analytical, P3/RT3) through the
drop-in wave_2D_solve call on one GPU; the CPU oracle (sparse LU, factor once) beside it on a bounded number of steps.
Run: python profiles/ho_cases.py  (prints one JSON line per case)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from porthamiltonian_fem_b200 import wave_2D_solve
from porthamiltonian_fem_b200.mesh import rectangle_structured
from oracle.wave_2D_oracle import wave_2D_solve_oracle

K, RHO = 3.0, 2.0
steps = int(os.environ.get('HO_STEPS', '3000'))
cpu_steps = int(os.environ.get('CPU_STEPS', '30'))
CASES = [('SV', (3, 3), (136, 34)), ('SE', (3, 3), (136, 34)), ('SV', (2, 2), (136, 34)), ('SV', (1, 2), (34, 8))]
if os.environ.get('HO_ONLY'):          # one case only (ncu captures)
    CASES = [CASES[int(os.environ['HO_ONLY'])]]
for scheme, bo, (Nx, Ny) in CASES:
    mesh = rectangle_structured(Nx, Ny)
    tF = 1.5
    t0 = time.time()
    H, nc, st = wave_2D_solve(tF, steps, '/tmp', 0, 1.0, 0.25, K_wave=K, rho=RHO, mesh=mesh, timeIntScheme=scheme, analytical=True,
                              basis_order=bo, return_state=True, verbose=False)
    wall = time.time() - t0
    stats = st['stats']
    ndof = len(st['p']) + len(st['q'])
    timing = {}
    t1 = time.time()
    Ho, _ = wave_2D_solve_oracle(tF * cpu_steps / steps, cpu_steps, mesh, 1.0, 0.25, K_wave=K, rho=RHO, timeIntScheme=scheme,
                                 analytical=True, basis_order=bo, timing=timing)
    err = np.abs(H[:cpu_steps, 1:] - Ho[:, 1:]).max() / np.abs(Ho[:, 1:]).max()
    print(json.dumps({'case': 'R {} weak {}x{}x2 analytical P{}/RT{}'.format(scheme, Nx, Ny, *bo), 'cells': nc, 'dofs': ndof, 'steps': steps,
                      'gpu_device_ms': stats['device_ms'], 'gpu_steps_per_s': steps / (stats['device_ms'] * 1e-3),
                      'gpu_wall_s_incl_setup': wall, 'launches': stats['kernel_launches'], 'matrix_free': bool(stats['kernel_paths'] & 1024),
                      'iters_p': stats['iters_p'] / max(1, stats['solves_p']), 'iters_q': stats['iters_q'] / max(1, stats['solves_q']),
                      'cpu_oracle_steps_per_s': cpu_steps / timing['loop_seconds'], 'cpu_oracle_total_s_incl_factorisation': time.time() - t1,
                      'H_array_rel_diff_first_%d_steps' % cpu_steps: err, 'errL2_end': float(H[-1, 8]), 'max_abs_E': float(np.abs(H[:, 2]).max())}))
    sys.stdout.flush()
