"""Stand-alone 4-state lumped model (electroMechanical.py of the reference): the restated Stormer-Verlet / symplectic
Euler steppers against scipy's solve_ivp on the same ODE (the check electromechanical_multi_nonFenics.py:217 makes in the
reference), their convergence orders, and the energy balance."""
import numpy as np
import scipy.integrate as si

from porthamiltonian_fem_b200.electroMechanical import electromechanical_solve, lumped_matrix_4, default_input


def reference_solution(tF):
    A, p = lumped_matrix_4()
    f = lambda t, x: A @ x + np.array([default_input(t), 0, 0, 0])
    sol = si.solve_ivp(f, (0, tF), np.zeros(4), rtol=1e-11, atol=1e-13, max_step=1e-2)
    return sol.y[:, -1], p


def test_schemes_converge_to_the_ode_solution():
    """forced run from rest: every scheme converges to the ODE solution at first order (the reference evaluates the input at
    t_{n+1} in all sub-steps, electroMechanical.py:300-301); Stormer-Verlet has the smallest constant"""
    tF = 2.0
    xe, p = reference_solution(tF)
    He = 0.5 * (xe[0] ** 2 / p['L_em'] + xe[1] ** 2 / p['m_em'] + xe[2] ** 2 / p['C_em'] + p['K_em'] * xe[3] ** 2)
    errs = {}
    for scheme in ('SV', 'SE', 'IE', 'EE'):
        errs[scheme] = []
        for n in (8000, 16000):
            H, E, t, d = electromechanical_solve(tF, n, timeIntScheme=scheme)
            assert len(H) == n + 1 and abs(t[-1] - tF) < 1e-9
            errs[scheme].append(abs(d[-1] - xe[3]) + abs(H[-1] - He))
        assert errs[scheme][1] < 1e-4 and 1.5 < errs[scheme][0] / errs[scheme][1] < 3.0, (scheme, errs[scheme])
    assert errs['SV'][1] < 0.2 * min(errs['SE'][1], errs['IE'][1])


def test_free_oscillation_orders():
    """no input: Stormer-Verlet is second order, symplectic Euler first order; both conserve the Hamiltonian up to a bounded
    oscillation, explicit / implicit Euler gain / lose energy"""
    A, p = lumped_matrix_4()
    x0 = np.array([0.01, 0.0, 0.0, 0.02])
    tF = 1.0
    sol = si.solve_ivp(lambda t, x: A @ x, (0, tF), x0, rtol=1e-12, atol=1e-14)
    zero = lambda t: 0.0
    err = {}
    for scheme in ('SV', 'SE'):
        err[scheme] = [abs(electromechanical_solve(tF, n, timeIntScheme=scheme, u_input=zero, x_init=x0)[3][-1] - sol.y[3, -1]) for n in (2000, 4000)]
    assert 3.5 < err['SV'][0] / err['SV'][1] < 4.5 and 1.7 < err['SE'][0] / err['SE'][1] < 2.3
    H0 = 0.5 * (x0[0] ** 2 / p['L_em'] + p['K_em'] * x0[3] ** 2)
    Hsv = np.array(electromechanical_solve(20.0, 20000, timeIntScheme='SV', u_input=zero, x_init=x0)[0][1:])
    Hee = np.array(electromechanical_solve(20.0, 20000, timeIntScheme='EE', u_input=zero, x_init=x0)[0][1:])
    Hie = np.array(electromechanical_solve(20.0, 20000, timeIntScheme='IE', u_input=zero, x_init=x0)[0][1:])
    assert np.abs(Hsv - H0).max() < 1e-3 * H0 and Hee[-1] > 1.5 * H0 and Hie[-1] < 0.7 * H0


def test_energy_balance_of_the_symplectic_schemes():
    """E = H_em - input energy stays at the level of the splitting error (no resistances: the model is lossless)"""
    H, E, t, d = electromechanical_solve(6.0, 6000, timeIntScheme='SV')
    assert np.abs(E).max() < 1e-5 * max(H)
    H1, E1, _, _ = electromechanical_solve(6.0, 6000, timeIntScheme='SE')
    assert np.abs(E1).max() < 5e-2 * max(H1)
    # after the input stops (t > 5) the Hamiltonian of the Stormer-Verlet run is conserved to the same level
    Hs = np.array(H)[np.array(t) > 5.0]
    assert np.ptp(Hs) < 1e-4 * Hs.mean()
