"""Pins the CPU oracle against the reference's own golden traces (tests/golden/reference_H_arrays.npz,
extracted from fenics_projects/wave_eqn/test_output/*/ground_truth by tests/golden/make_golden.py; these are
the arrays the reference's test_case_check.py:80-112 compares against).

The reference meshes come from mshr/CGAL and cannot be rebuilt here, so:
  * exact checks: the t column, and every entry that does not depend on the mesh (first-step lumped
    quantities of the IC cases, first-step nodal error of the analytical cases) - 11 to 13 digits;
  * discretisation-level checks on a structured mesh of comparable resolution: Hamiltonian / displacement
    traces within a few 1e-4 (IC) or a few percent (coarse wave cases), energy-residual magnitude and sign;
  * the scratch known-answer values of SURVEY.md Appendix B (same mesh family) to 10 digits.
"""
import os

import numpy as np
import pytest

from oracle.wave_2D_oracle import wave_2D_solve_oracle
from porthamiltonian_fem_b200.mesh import rectangle_structured, generate_mesh

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'reference_H_arrays.npz'))
K_WAVE, RHO = 3.0, 2.0          # run_test_cases.py:18-20


def golden(case, group='test_cases'):
    key = group + '/' + case
    return G[key + '/rows'], G[key + '/H'], float(G[key + '/maxabsE'][0]), int(G[key + '/numSteps'][0])


def run(Nx, Ny, steps, tF, **kw):
    H, nc = wave_2D_solve_oracle(tF, steps, rectangle_structured(Nx, Ny, 1.0, 0.25), 1.0, 0.25, K_wave=K_WAVE, rho=RHO, **kw)
    return H


def test_golden_time_columns_exact():
    for key in [k[:-2] for k in G.files if k.endswith('/H')]:
        rows, H, n = G[key + '/rows'], G[key + '/H'], int(G[key + '/numSteps'][0])
        case = key.split('/')[1]
        tF = {'t0_1': 0.1, 't1_5': 1.5, 't4_5': 4.5, 't8': 8.0}[[s for s in ('t0_1', 't1_5', 't4_5', 't8') if ('_' + s + '_') in case][0]]
        dt = tF / n
        t = 0.0
        acc = []
        for i in range(n):
            t += dt                                   # the reference accumulates t (wave_2D.py:791,834)
            acc.append(t)
        assert np.array_equal(np.array(acc)[rows], H[:, 0]), key


def test_ic_sv_first_step_exact():
    """Row 0 of every IC/SV golden trace is a pure lumped-ODE step (wave state still zero): mesh independent."""
    for group, case, tF, n in [('test_cases', 'R_SV_weak_nx60_IC_t0_1_steps100', 0.1, 100),
                               ('test_cases', 'S_1C_SV_weak_nx40_IC_t0_1_steps100', 0.1, 100),
                               ('test_cases_long', 'R_SV_weak_nx120_IC_t4_5_steps3000', 4.5, 3000),
                               ('test_cases_long', 'S_1C_SV_weak_nx60_IC_t8_steps16000', 8.0, 16000)]:
        rows, Hg, _, _ = golden(case, group)
        H, _ = wave_2D_solve_oracle(tF / n * 2, 2, rectangle_structured(8, 2), 1.0, 0.25, K_wave=K_WAVE, rho=RHO,
                                    timeIntScheme='SV', interConnection='IC')
        for col in (3, 5, 6):   # disp, inputEnergy, H_em
            assert abs(H[0, col] - Hg[0, col]) <= 2e-13 * abs(Hg[0, col]), (case, col, H[0, col], Hg[0, col])


def test_ic_sm_first_step_lumped_energy():
    rows, Hg, _, _ = golden('R_SM_weak_nx60_IC_t0_1_steps100')
    H = run(104, 26, 100, 0.1, timeIntScheme='SM', interConnection='IC')
    assert abs(H[0, 6] - Hg[0, 6]) < 1e-6 * abs(Hg[0, 6])      # H_em, coupling to the wave is O(1e-14)
    assert abs(H[0, 3] - Hg[0, 3]) < 1e-3 * abs(Hg[0, 3])


def test_analytical_first_step_nodal_error():
    """err_max after one SE step is attained at a boundary node and is mesh independent to ~1e-11."""
    rows, Hg, _, _ = golden('R_SE_weak_nx60_analytical_t0_1_steps200_P1_RT1')
    H = run(104, 26, 200, 0.1, timeIntScheme='SE', analytical=True)
    assert abs(H[0, 9] - Hg[0, 9]) < 1e-9 * Hg[0, 9]
    # discretisation-level: Hamiltonian trace and point values
    assert np.max(np.abs(H[rows, 1] - Hg[:, 1]) / Hg[:, 1]) < 2e-3
    assert abs(H[-1, 2] - Hg[-1, 2]) < 0.1 * abs(Hg[-1, 2])
    assert abs(H[0, 11] - Hg[0, 11]) < 2e-2 * abs(Hg[0, 11])


@pytest.mark.parametrize('case,Nx,Ny,kw,tolH', [
    ('R_SV_weak_nx60_IC_t0_1_steps100', 104, 26, dict(timeIntScheme='SV', interConnection='IC'), 5e-4),
    ('R_SM_weak_nx60_IC_t0_1_steps100', 104, 26, dict(timeIntScheme='SM', interConnection='IC'), 5e-4),
])
def test_ic_traces_discretisation_level(case, Nx, Ny, kw, tolH):
    rows, Hg, maxE, n = golden(case)
    H = run(Nx, Ny, n, 0.1, **kw)
    late = rows[rows > 20]
    for col in (1, 3, 6):      # H, disp, H_em
        rel = np.abs(H[late, col] - Hg[late, col]) / np.abs(Hg[late, col]).max()
        assert rel.max() < tolH, (case, col, rel.max())
    if kw['timeIntScheme'] == 'SM':
        assert np.abs(H[:, 2]).max() < 1e-13 and maxE < 1e-13      # implicit midpoint conserves the energy balance
    else:
        assert abs(np.abs(H[:, 2]).max() - maxE) < 0.01 * maxE      # SV residual ~ dt^2, golden 4.558e-06


@pytest.mark.parametrize('case,Nx,Ny,scheme,sign', [
    ('R_EE_weak_nx40_wave_t0_1_steps200', 68, 17, 'EE', +1),
    ('R_IE_weak_nx40_wave_t0_1_steps200', 68, 17, 'IE', -1),
    ('R_EH_weak_nx40_wave_t0_1_steps200', 68, 17, 'EH', +1),
    ('R_SV_weak_nx20_wave_t0_1_steps200', 34, 8, 'SV', +1),
])
def test_wave_traces_discretisation_level(case, Nx, Ny, scheme, sign):
    rows, Hg, maxE, n = golden(case)
    H = run(Nx, Ny, n, 0.1, timeIntScheme=scheme)
    assert abs(H[-1, 1] - Hg[-1, 1]) < 0.03 * Hg[-1, 1]
    assert np.sign(H[-1, 2]) == sign == np.sign(Hg[-1, 2])
    assert 0.6 < H[-1, 2] / Hg[-1, 2] < 1.4
    assert abs(H[-1, 4] - Hg[-1, 4]) < 0.03 * abs(Hg[-1, 4])


def test_strong_first_step():
    rows, Hg, _, n = golden('R_SV_strong_nx20_wave_t0_1_steps200')
    H = run(34, 8, n, 0.1, timeIntScheme='SV', dirichletImp='strong')
    assert abs(H[0, 1] - Hg[0, 1]) < 0.03 * Hg[0, 1]
    assert abs(H[-1, 1] - Hg[-1, 1]) < 0.03 * Hg[-1, 1]


def test_appendix_b_known_answers():
    """SURVEY.md Appendix B values (structured mesh, exact sparse solves)."""
    H = run(34, 8, 200, 0.1, timeIntScheme='SV')
    assert abs(H[-1, 1] - 0.91118461659) < 1e-10 and abs(H[-1, 2] - 1.1473722557e-05) < 1e-14
    H = run(68, 17, 200, 0.1, timeIntScheme='EH')
    assert abs(H[-1, 1] - 0.90951467175) < 1e-10 and abs(H[-1, 2] - 5.7754852781e-06) < 1e-14
    H = run(104, 26, 100, 0.1, timeIntScheme='SV', interConnection='IC')
    assert abs(H[-1, 1] - 0.25722725041) < 1e-10 and abs(H[-1, 3] + 0.072615625976) < 1e-11
    assert abs(H[0, 3] + 1.046651205470116e-08) < 1e-20


def test_long_run_energy_balance_sm():
    """R_SM_weak_nx120_IC long golden: max|E| ~ 7e-13 over 3000 steps; oracle on a coarser mesh, 300 steps."""
    _, _, maxE, _ = golden('R_SM_weak_nx120_IC_t4_5_steps3000', 'test_cases_long')
    H = run(52, 13, 300, 0.45, timeIntScheme='SM', interConnection='IC')
    assert np.abs(H[:, 2]).max() < 1e-12 and maxE < 1e-11


def test_s1c_trace():
    rows, Hg, maxE, n = golden('S_1C_SV_weak_nx40_IC_t0_1_steps100')
    mesh = generate_mesh('S_1C', 40, 1.0, 0.1)
    H, nc = wave_2D_solve_oracle(0.1, n, mesh, 1.0, 0.1, domainShape='S_1C', K_wave=K_WAVE, rho=RHO,
                                 timeIntScheme='SV', interConnection='IC')
    assert abs(H[-1, 1] - Hg[-1, 1]) < 5e-3 * Hg[-1, 1]
    assert abs(H[-1, 3] - Hg[-1, 3]) < 5e-3 * abs(Hg[-1, 3])
    assert abs(H[-1, 6] - Hg[-1, 6]) < 5e-3 * abs(Hg[-1, 6])


def test_invalid_combinations_quit():
    mesh = rectangle_structured(4, 2)
    for kw in [dict(controlType='lqr'), dict(controlType='passivity'), dict(controlType='casimir', interConnection='IC'),
               dict(interConnection='IC', dirichletImp='strong'), dict(interConnection='IC', timeIntScheme='EH'),
               dict(timeIntScheme='XX')]:
        with pytest.raises(SystemExit):
            wave_2D_solve_oracle(0.1, 2, mesh, 1.0, 0.25, **kw)
