"""Regression harness: the case lists of the reference's run_test_cases.py (:23-49, all eleven tuples of each set, the
higher-order analytical ones included) on the B200 path, written to
test_output/<caseName>/..., then compared by test_case_check.check_cases against a ground truth produced by the CPU
oracle on the same meshes (the reference's own ground-truth files were produced on mshr meshes that cannot be
rebuilt here; they pin the oracle instead - see tests/test_oracle_golden.py).

    python run_test_cases.py [test_cases|test_cases_long]
"""
import os
import sys
import time

import numpy as np

import test_case_check
from main_wave_2D import run_cases, sub_dir_name, K_wave, rho

CASE_SETS = {
    'test_cases': [['R', 'IE', 'weak', 40, 'wave', 0.1, 200, 'noMem'],
                   ['R', 'EE', 'weak', 40, 'wave', 0.1, 200, 'noMem'],
                   ['R', 'EH', 'weak', 40, 'wave', 0.1, 200, 'noMem'],
                   ['R', 'SV', 'strong', 20, 'wave', 0.1, 200, 'noMem'],
                   ['R', 'SV', 'weak', 20, 'wave', 0.1, 200, 'noMem'],
                   ['R', 'SV', 'weak', 4, 'analytical', 0.1, 200, 'noMem', (3, 3)],
                   ['R', 'SV', 'weak', 20, 'analytical', 0.1, 200, 'noMem', (1, 2)],
                   ['R', 'SE', 'weak', 60, 'analytical', 0.1, 200, 'noMem', (1, 1)],
                   ['R', 'SV', 'weak', 60, 'IC', 0.1, 100, 'noMem'],
                   ['R', 'SM', 'weak', 60, 'IC', 0.1, 100, 'noMem'],
                   ['S_1C', 'SV', 'weak', 40, 'IC', 0.1, 100, 'noMem']],
    'test_cases_long': [['R', 'IE', 'weak', 80, 'wave', 1.5, 3000, 'noMem'],
                        ['R', 'EE', 'weak', 80, 'wave', 1.5, 3000, 'noMem'],
                        ['R', 'EH', 'weak', 80, 'wave', 1.5, 3000, 'noMem'],
                        ['R', 'SV', 'strong', 20, 'wave', 1.5, 3000, 'noMem'],
                        ['R', 'SV', 'weak', 20, 'wave', 1.5, 3000, 'noMem'],
                        ['R', 'SV', 'weak', 4, 'analytical', 1.5, 3000, 'noMem', (3, 3)],
                        ['R', 'SV', 'weak', 20, 'analytical', 1.5, 3000, 'noMem', (1, 2)],
                        ['R', 'SE', 'weak', 80, 'analytical', 1.5, 3000, 'noMem', (3, 3)],
                        ['R', 'SV', 'weak', 120, 'IC', 4.5, 3000, 'noMem'],
                        ['R', 'SM', 'weak', 120, 'IC', 4.5, 3000, 'noMem'],
                        ['S_1C', 'SV', 'weak', 60, 'IC', 8, 16000, 'noMem']],
}


def make_ground_truth(caseName, caseArray):
    """CPU oracle on the same meshes -> test_output/<caseName>/ground_truth/<subDir>/H_array.npy"""
    from oracle.wave_2D_oracle import wave_2D_solve_oracle
    from porthamiltonian_fem_b200.mesh import generate_mesh
    for caseVec in caseArray:
        yLength = 0.25 if caseVec[0] == 'R' else 0.1
        d = os.path.join('test_output', caseName, 'ground_truth', sub_dir_name(caseVec))
        os.makedirs(d, exist_ok=True)
        mesh = generate_mesh(caseVec[0], caseVec[3], 1.0, yLength)
        H, nc = wave_2D_solve_oracle(caseVec[5], caseVec[6], mesh, 1.0, yLength, domainShape=caseVec[0],
                                     timeIntScheme=caseVec[1], dirichletImp=caseVec[2], K_wave=K_wave, rho=rho,
                                     interConnection=caseVec[4] if caseVec[4].startswith('IC') else None,
                                     analytical=caseVec[4] == 'analytical',
                                     basis_order=caseVec[8] if len(caseVec) > 8 else (1, 1))
        np.save(os.path.join(d, 'H_array.npy'), H)
        np.save(os.path.join(d, 'numCells.npy'), np.array([float(nc)]))


if __name__ == '__main__':
    print('starting script and timer')
    tic = time.time()
    caseName = sys.argv[1] if len(sys.argv) > 1 else 'test_cases'
    caseArray = CASE_SETS[caseName]
    run_cases(caseName, caseArray, outputRoot='test_output', verbose=False)
    print('All Simulations finished in {} seconds'.format(time.time() - tic))
    print('now running test case check')
    make_ground_truth(caseName, caseArray)
    ok = test_case_check.check_cases(caseName, caseArray)
    sys.exit(0 if ok else 1)
