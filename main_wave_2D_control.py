"""Control driver: same case tuples and outputs as the reference's main_wave_2D_control.py (:31-41 cases,
:89-98 controlType / casimir xLength, :43-48 output tree), running on the B200 path.

    python main_wave_2D_control.py [passivity|casimir] [quick]
"""
import os
import sys
import time

import numpy as np

from main_wave_2D import sub_dir_name
from porthamiltonian_fem_b200 import wave_2D_solve

K_wave = 3.0
rho = 2.0
SAVE_EVERY = int(os.environ.get('PHW_SAVE_EVERY', '1'))   # 'Mem' cases: p.xdmf at every step as in the reference (wave_2D.py:865-866)

if __name__ == '__main__':
    print('starting script and timer')
    tic = time.time()
    which = sys.argv[1] if len(sys.argv) > 1 else 'casimir'
    quick = len(sys.argv) > 2 and sys.argv[2] == 'quick'
    if which == 'casimir':
        caseName = 'casimir_control_wave_3_0'
        caseArray = [['R', 'SM', 'weak', 5, 'wave', 0.4, 400, 'Mem', 'casimir']] if quick else \
                    [['R', 'SM', 'weak', 100, 'wave', 5.0, 5000, 'Mem', 'casimir']]
    else:
        caseName = 'passivity_control_IC_t4_5'
        caseArray = [['R', 'SM', 'weak', 40, 'IC', 0.4, 400, 'Mem', 'passivity']] if quick else \
                    [['R', 'SM', 'weak', 40, 'IC', 8.0, 8000, 'Mem', 'passivity']]
    input_stop_t = 0.0
    control_start_t = 0.2 if quick else 1.0
    outputDir = os.path.join('control_output', caseName)
    os.makedirs(outputDir, exist_ok=True)
    for caseVec in caseArray:
        IC = caseVec[4] if caseVec[4].startswith('IC') else None
        controlType = caseVec[8] if len(caseVec) > 8 else None
        outputSubDir = os.path.join(outputDir, sub_dir_name(caseVec[:8]))
        os.makedirs(outputSubDir, exist_ok=True)
        xLength = 4.0 if controlType == 'casimir' else 1.0
        yLength = 0.25 if caseVec[0] == 'R' else 0.1
        H_array, numCells = wave_2D_solve(caseVec[5], caseVec[6], outputSubDir, caseVec[3], xLength, yLength,
                                          domainShape=caseVec[0], timeIntScheme=caseVec[1], dirichletImp=caseVec[2],
                                          K_wave=K_wave, rho=rho, interConnection=IC, analytical=False,
                                          saveP=caseVec[7] != 'noMem', save_every=SAVE_EVERY if caseVec[7] != 'noMem' else None, controlType=controlType,
                                          input_stop_t=input_stop_t, control_start_t=control_start_t)
        np.save(os.path.join(outputSubDir, 'H_array.npy'), H_array)
        numCells_save = np.zeros((1))
        numCells_save[0] = numCells
        np.save(os.path.join(outputSubDir, 'numCells.npy'), numCells_save)
    print('All Simulations finished in {} seconds'.format(time.time() - tic))
