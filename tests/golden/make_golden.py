"""Extracts the reference's own golden traces (fenics_projects/wave_eqn/test_output/*/ground_truth/*/H_array.npy,
checked by the reference's test_case_check.py:80-112) into one small fixture file.

Run in the build container (where /root/reference is mounted):  python tests/golden/make_golden.py
Quick-set traces are stored in full; long-set traces as rows {0,1,2, every 100th, last} plus max|E|.
"""
import glob
import os
import numpy as np

REF = '/root/reference/fenics_projects/wave_eqn/test_output'
out = {}
for d in sorted(glob.glob(os.path.join(REF, '*', 'ground_truth', '*'))):
    case = os.path.basename(d)
    group = os.path.basename(os.path.dirname(os.path.dirname(d)))
    H = np.load(os.path.join(d, 'H_array.npy'))
    nc = np.load(os.path.join(d, 'numCells.npy'))
    key = group + '/' + case
    if group == 'test_cases':
        rows = np.arange(H.shape[0])
    else:
        rows = np.unique(np.concatenate([[0, 1, 2], np.arange(0, H.shape[0], 100), [H.shape[0] - 1]]))
    out[key + '/rows'] = rows.astype(np.int64)
    out[key + '/H'] = H[rows]
    out[key + '/numCells'] = nc
    out[key + '/numSteps'] = np.array([H.shape[0]])
    out[key + '/maxabsE'] = np.array([np.abs(H[:, 2]).max()])
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'reference_H_arrays.npz'), **out)
print('wrote', len(out) // 5, 'cases')
