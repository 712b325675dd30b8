"""Eigenvalue driver: same case tuples and outputs as the reference's main_eigen.py (:17-23 case array, P2 x RT2 on nx = 5 .. 25; out
array columns exact / model / real part), running the device operators (porthamiltonian_fem_b200/calculate_eigen.py).

    python main_eigen.py
"""
import os
import time

import numpy as np

from main_wave_2D import sub_dir_name
from porthamiltonian_fem_b200.calculate_eigen import calculate_eigen

K_wave = 3.0
rho = 2.0

if __name__ == '__main__':
    print('starting script and timer')
    tic = time.time()
    caseName = 'eigenValue_spaceVariation'
    caseArray = [['R', 'SE', 'weak', nx, 'analytical', 0.5, 1000, 'noMem', (2, 2)] for nx in (5, 10, 15, 20, 25)]      # main_eigen.py:17-23
    outputDir = os.path.join('output', caseName)
    for caseVec in caseArray:
        outputSubDir = os.path.join(outputDir, sub_dir_name(caseVec))
        os.makedirs(outputSubDir, exist_ok=True)
        out_array, numCells = calculate_eigen(caseVec[5], caseVec[6], outputSubDir, caseVec[3], 1.0, 0.25, domainShape=caseVec[0],
                                              timeIntScheme=caseVec[1], dirichletImp=caseVec[2], K_wave=K_wave, rho=rho,
                                              analytical=True, basis_order=caseVec[8])
        np.save(os.path.join(outputSubDir, 'eigen_array.npy'), out_array)
        numCells_save = np.zeros((1))
        numCells_save[0] = numCells
        np.save(os.path.join(outputSubDir, 'numCells.npy'), numCells_save)
        print('nx = {}: {} cells, first eigenvalues exact / model'.format(caseVec[3], numCells))
        print(out_array[:6, :2])
    print('All eigenvalue cases finished in {} seconds'.format(time.time() - tic))
