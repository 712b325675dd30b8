"""Driver of the stand-alone 4-state lumped model: the reference's main_electroMechanical.py (Stormer-Verlet and symplectic Euler runs of
electromechanical_solve, saved as H / E / t / disp arrays).

    python main_electroMechanical.py
"""
import os

import numpy as np

from porthamiltonian_fem_b200.electroMechanical import electromechanical_solve

if __name__ == '__main__':
    outputDir = os.path.join('output', 'electromechanical')
    os.makedirs(outputDir, exist_ok=True)
    for scheme in ('SV', 'SE'):
        H, E, t, d = electromechanical_solve(10.0, 10000, outputDir, 20, 5, 1.0, 0.25, timeIntScheme=scheme)
        np.save(os.path.join(outputDir, 'HEtd_{}.npy'.format(scheme)), np.array([H, E, t, d]))
        print('{}: max |E| = {:.3e}, H_end = {:.6e}, disp_end = {:.6e}'.format(scheme, np.abs(E).max(), H[-1], d[-1]))
