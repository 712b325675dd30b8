"""Higher-order pairs on meshes whose assembled operators do not fit the 126 MB L2: matrix-free cell kernels (56 B per cell) against
assembled CSR rows (12 B per non-zero).  Decides the size above which wave_2D_solve takes the matrix-free path by default.
Run: python profiles/ho_large.py  (one JSON line per case and path)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from porthamiltonian_fem_b200 import wave_2D_solve
from porthamiltonian_fem_b200.mesh import rectangle_structured

K, RHO = 3.0, 2.0
steps = int(os.environ.get('HO_STEPS', '40'))
for bo, (Nx, Ny) in [((2, 2), (816, 204)), ((3, 3), (408, 102)), ((2, 2), (272, 68))]:
    mesh = rectangle_structured(Nx, Ny)
    for csr in (False, True):
        os.environ.pop('PHW_HO_CSR', None)
        os.environ.pop('PHW_HO_MF', None)
        os.environ['PHW_HO_CSR' if csr else 'PHW_HO_MF'] = '1'
        t0 = time.time()
        H, nc, st = wave_2D_solve(1.5 * steps / 30000, steps, '/tmp', 0, 1.0, 0.25, K_wave=K, rho=RHO, mesh=mesh, timeIntScheme='SV', analytical=True,
                                  basis_order=bo, return_state=True, verbose=False)
        stats = st['stats']
        print(json.dumps({'case': 'R SV weak {}x{}x2 analytical P{}/RT{}'.format(Nx, Ny, *bo), 'cells': nc, 'dofs': len(st['p']) + len(st['q']), 'steps': steps,
                          'gpu_device_ms': stats['device_ms'], 'gpu_steps_per_s': steps / (stats['device_ms'] * 1e-3),
                          'gpu_wall_s_incl_setup': time.time() - t0, 'matrix_free': bool(stats['kernel_paths'] & 1024),
                          'iters_p': stats['iters_p'] / max(1, stats['solves_p']), 'iters_q': stats['iters_q'] / max(1, stats['solves_q']),
                          'H_end': float(H[-1, 1]), 'errL2_end': float(H[-1, 8])}), flush=True)
