"""Stand-alone 4-state lumped electromechanical model (reference: electroMechanical.py:11-371, driver
main_electroMechanical.py): the port-Hamiltonian ODE  x_dot = (J - R) dH/dx + e_0 u(t)  with x = (flux linkage, momentum,
capacitor charge, displacement), H = 1/2 (x0^2/L + x1^2/m + x2^2/C + K x3^2) (:56-70), input u(t) = sin(pi t) for t < 5 (:152),
integrated with the reference's Stormer-Verlet or symplectic-Euler splitting (x0, x1 against x2, x3; forms :160-199; the input
is evaluated at t_{n+1} in every sub-step, :300-301, so a forced run is first-order accurate in dt with every scheme).

The reference poses it as a FEniCS problem on four 'R' (real) spaces, i.e. a 4 x 4 linear solve per step with no mesh dependence;
here it is exactly that 4 x 4 system in numpy.  It is the cross-check of the lumped-model stepper: the 3-state model that the wave
solver couples to (wave_2D.py:148-166, integrated on the device by scalar_k) is this model without the capacitor.  The explicit /
implicit Euler variants of the reference (:201-213) multiply the 4 x 4 matrix by a 3-vector and cannot run there either; they are
restated with the full state.
"""
import numpy as np

DEFAULTS = dict(Bl_em=5.0, R1_em=0.0, R2_em=0.0, K_em=5.0, m_em=0.15, L_em=0.1, C_em=1.0)     # electroMechanical.py:56-62


def lumped_matrix_4(prm=None):
    """A_em = (J - R) diag(1/L, 1/m, 1/C, K)   (electroMechanical.py:64-70)"""
    p = dict(DEFAULTS)
    if prm:
        p.update(prm)
    J = np.array([[0, p['Bl_em'], -1, 0], [-p['Bl_em'], 0, 0, -1], [1, 0, 0, 0], [0, 1, 0, 0]], dtype=float)
    R = np.diag([p['R1_em'], p['R2_em'], 0.0, 0.0])
    return (J - R) @ np.diag([1 / p['L_em'], 1 / p['m_em'], 1 / p['C_em'], p['K_em']]), p


def default_input(t):
    return np.sin(np.pi * t) if t < 5.0 else 0.0                                              # :152


def electromechanical_solve(tFinal, numSteps, outputDir=None, nx=None, ny=None, xLength=None, yLength=None, domainShape='R',
                            timeIntScheme='SV', dirichletImp='weak', lumped=None, u_input=default_input, x_init=None):
    """Same call and return value as the reference (electroMechanical.py:11-13,371): (H_vec, E_vec, t_vec, disp_vec), lists that start
    with the t = 0 entry.  E = H_em - input energy (:343)."""
    if dirichletImp != 'weak':
        print('can only do interconnection model with weak boundary conditions')
        raise SystemExit
    A, p = lumped_matrix_4(lumped)
    dt = tFinal / numSteps
    I = np.eye(4)
    a, b = [0, 1], [2, 3]                      # the two halves of the splitting
    x = np.zeros(4) if x_init is None else np.array(x_init, dtype=float)      # the reference starts from rest
    H_vec, E_vec, t_vec, disp_vec = [0], [0], [0], [0]
    inputEnergy = 0.0
    t = 0.0
    if timeIntScheme == 'SV':
        # stage 1 (F_temp, :160-171): half step, x_a old, x_b implicit, input on row 0.   stage 2 (F, :173-189): full step on x_a with
        # the trapezoidal rule in x_a and x_b = stage value; x_b advances another half step from the stage value with the new x_a
        M1 = I.copy()
        M1[:, b] -= 0.5 * dt * A[:, b]
        M2 = I.copy()
        M2[np.ix_(a, a)] -= 0.5 * dt * A[np.ix_(a, a)]
        M2[np.ix_(b, a)] -= 0.5 * dt * A[np.ix_(b, a)]
    elif timeIntScheme == 'SE':
        M = I.copy()
        M[:, a] -= dt * A[:, a]                # x_a implicit, x_b old (:191-199)
    elif timeIntScheme == 'IE':
        M = I - dt * A
    elif timeIntScheme != 'EE':
        print('time integration scheme {} not implemented'.format(timeIntScheme))
        raise SystemExit
    for n in range(numSteps):
        t += dt                                                                                 # :300
        u = u_input(t)
        xn = x.copy()
        if timeIntScheme == 'SV':
            r1 = xn + 0.5 * dt * (A[:, a] @ xn[a])
            r1[0] += 0.5 * dt * u
            xt = np.linalg.solve(M1, r1)
            r2 = np.zeros(4)
            r2[a] = xn[a] + dt * (0.5 * A[np.ix_(a, a)] @ xn[a] + A[np.ix_(a, b)] @ xt[b])
            r2[0] += dt * u                                                                     # full-step input (:189)
            r2[b] = xt[b] + 0.5 * dt * (A[np.ix_(b, b)] @ xt[b])
            x = np.linalg.solve(M2, r2)
            inputEnergy += 0.5 * dt * u * (xn[0] + x[0]) / p['L_em']                            # :317
        elif timeIntScheme == 'SE':
            r = xn + dt * (A[:, b] @ xn[b])
            r[0] += dt * u
            x = np.linalg.solve(M, r)
            inputEnergy += dt * u * x[0] / p['L_em']                                            # :319
        elif timeIntScheme == 'IE':
            r = xn.copy()
            r[0] += dt * u
            x = np.linalg.solve(M, r)
            inputEnergy += dt * u * x[0] / p['L_em']                                            # (:318 'SE' or 'IE' is always true)
        else:   # EE
            x = xn + dt * (A @ xn)
            x[0] += dt * u
            inputEnergy += dt * u * x[0] / p['L_em']
        H_em = 0.5 * (x[0] ** 2 / p['L_em'] + x[1] ** 2 / p['m_em'] + x[2] ** 2 / p['C_em'] + p['K_em'] * x[3] ** 2)   # :331-332
        H_vec.append(H_em)
        E_vec.append(-inputEnergy + H_em)
        t_vec.append(t)
        disp_vec.append(x[3])
    return H_vec, E_vec, t_vec, disp_vec
