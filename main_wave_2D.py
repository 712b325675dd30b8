"""Paper-case driver: same case tuples, naming and outputs as the reference's main_wave_2D.py (:104-114 case
array, :153-182 sub-directory naming, :185-191 lengths, :214-219 outputs), running on the B200 path.

    python main_wave_2D.py [caseName]
"""
import os
import sys
import time

import numpy as np

from porthamiltonian_fem_b200 import wave_2D_solve

K_wave = 3.0   # main_wave_2D.py:19
rho = 2.0      # main_wave_2D.py:21
# 'Mem' cases write p.xdmf (and p_exact.xdmf) at every step as the reference does (wave_2D.py:865-866,964-965); PHW_SAVE_EVERY thins it
SAVE_KW = dict(save_every=int(os.environ.get('PHW_SAVE_EVERY', '1')))

CASES = {
    # main_wave_2D.py:104-114 (the active block of the reference: P3/RT3, SV and SE, five step counts each)
    'analytical_t1_5_timeVariation': [['R', 'SV', 'weak', 80, 'analytical', 1.5, 3000, 'Mem', (3, 3)],
                                      ['R', 'SV', 'weak', 80, 'analytical', 1.5, 4250, 'noMem', (3, 3)],
                                      ['R', 'SV', 'weak', 80, 'analytical', 1.5, 3500, 'noMem', (3, 3)],
                                      ['R', 'SV', 'weak', 80, 'analytical', 1.5, 5000, 'noMem', (3, 3)],
                                      ['R', 'SV', 'weak', 80, 'analytical', 1.5, 6000, 'noMem', (3, 3)],
                                      ['R', 'SE', 'weak', 80, 'analytical', 1.5, 3000, 'noMem', (3, 3)],
                                      ['R', 'SE', 'weak', 80, 'analytical', 1.5, 3500, 'noMem', (3, 3)],
                                      ['R', 'SE', 'weak', 80, 'analytical', 1.5, 4250, 'noMem', (3, 3)],
                                      ['R', 'SE', 'weak', 80, 'analytical', 1.5, 5000, 'noMem', (3, 3)],
                                      ['R', 'SE', 'weak', 80, 'analytical', 1.5, 6000, 'noMem', (3, 3)]],
    # the same block at lowest order (P1/RT1)
    'analytical_t1_5_P1_RT1': [['R', 'SV', 'weak', 80, 'analytical', 1.5, 3000, 'Mem', (1, 1)],
                               ['R', 'SE', 'weak', 80, 'analytical', 1.5, 3000, 'noMem', (1, 1)]],
    # main_wave_2D.py:24-36
    'strong_and_weak_method_variation': [['R', s, 'weak', 80, 'wave', 1.5, 3000, 'Mem'] for s in ('IE', 'EE', 'SE', 'EH', 'SV', 'SM')],
    # main_wave_2D.py:92-102
    'IC_t4_5_timeVariation': [['R', 'SV', 'weak', 120, 'IC', 4.5, 3000, 'noMem'], ['R', 'SM', 'weak', 120, 'IC', 4.5, 3000, 'noMem']],
    # main_wave_2D.py:116-120
    'IC_SV_t_20_0_singleRun': [['R', 'SV', 'weak', 120, 'IC', 20, 40000, 'Mem']],
    'IC_square_IC_SV_t8_0': [['S_1C', 'SV', 'weak', 60, 'IC', 8, 16000, 'Mem']],
}


def sub_dir_name(caseVec):
    """main_wave_2D.py:157-175"""
    subDir = caseVec[0] + '_' + caseVec[1] + '_' + caseVec[2]
    subDir = subDir + '_nx' + str(caseVec[3])
    subDir = subDir + '_' + caseVec[4]
    subDir = subDir + '_t' + str(caseVec[5]).replace('.', '_') + '_steps' + str(caseVec[6])
    if len(caseVec) > 8:
        subDir = subDir + '_P{}_RT{}'.format(caseVec[8][0], caseVec[8][1])
    return subDir


def run_cases(caseName, caseArray, outputRoot='output', **solver_kwargs):
    results = []
    for caseVec in caseArray:
        assert len(caseVec) in [7, 8, 9], 'caseArray should be vectors of length 7, 8, or 9'
        outputDir = os.path.join(outputRoot, caseName)
        os.makedirs(outputDir, exist_ok=True)
        IC = caseVec[4] if caseVec[4].startswith('IC') else None
        ANALYTICAL_BOOL = (caseVec[4] == 'analytical')
        basis_order = caseVec[8] if len(caseVec) > 8 else (1, 1)
        outputSubDir = os.path.join(outputDir, sub_dir_name(caseVec))
        os.makedirs(outputSubDir, exist_ok=True)
        nx = caseVec[3]
        xLength = 1.0
        yLength = 0.25 if caseVec[0] == 'R' else 0.1
        saveP = not (len(caseVec) > 7 and caseVec[7] == 'noMem')
        H_array, numCells = wave_2D_solve(caseVec[5], caseVec[6], outputSubDir, nx, xLength, yLength,
                                          domainShape=caseVec[0], timeIntScheme=caseVec[1], dirichletImp=caseVec[2],
                                          K_wave=K_wave, rho=rho, interConnection=IC, analytical=ANALYTICAL_BOOL,
                                          saveP=saveP, basis_order=basis_order, **dict(SAVE_KW if saveP else {}, **solver_kwargs))
        np.save(os.path.join(outputSubDir, 'H_array.npy'), H_array)
        numCells_save = np.zeros((1))
        numCells_save[0] = numCells
        np.save(os.path.join(outputSubDir, 'numCells.npy'), numCells_save)
        results.append((outputSubDir, H_array, numCells))
    return results


if __name__ == '__main__':
    print('starting script and timer')
    tic = time.time()
    caseName = sys.argv[1] if len(sys.argv) > 1 else 'analytical_t1_5_timeVariation'   # main_wave_2D.py:104
    run_cases(caseName, CASES[caseName])
    print('All Simulations finished in {} seconds'.format(time.time() - tic))
