"""CPU oracle: numpy/scipy restatement of the reference's ``wave_2D_solve``.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Follows
``/root/reference/fenics_projects/wave_eqn/wave_2D.py`` region by region:

  argument checks ............ wave_2D.py:63-83,148-151      -> _check_case
  time levels per scheme ..... wave_2D.py:295-324            -> _ALPHA table (p1 = ap*p + (1-ap)*p_m, ...)
  boundary data .............. wave_2D.py:326-409            -> g_left(t), b_L weights, casimir blocks
  forms F / F_temp ........... wave_2D.py:449-526            -> _assemble_system (block matrices)
  lumped rows ................ wave_2D.py:148-166,528-577    -> the 3 trailing rows/cols
  A = assemble(lhs) once ..... wave_2D.py:579-595            -> scipy splu of the full mixed matrix
  initial conditions ......... wave_2D.py:732-771
  time loop .................. wave_2D.py:787-982            -> the for-loop below, same order of operations
  energies ................... wave_2D.py:600-641,876-927
  analytical diagnostics ..... wave_2D.py:929-965
  output packing ............. wave_2D.py:1011-1032

dolfin solves the *full* mixed system with a sparse LU at every step; the oracle does the same
with ``scipy.sparse.linalg.splu`` (factor once by default; ``refactor_every_solve=True`` mimics
dolfin's ``solve(A, x, b)``, which builds a fresh LU solver per call).

The mesh is an input (vertices, cells): mshr is not reproducible here.
"""
import time
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import fem

# (alpha_p, alpha_q): weight of the *new* value in p1 / q1 of wave_2D.py:295-324 (one-sub-step schemes)
_ALPHA = {'EE': (0.0, 0.0), 'IE': (1.0, 1.0), 'SE': (1.0, 0.0), 'SM': (0.5, 0.5)}

DEFAULT_LUMPED = dict(Bl_em=5.0, R1_em=0.0, R2_em=0.0, K_em=5.0, m_em=0.15, L_em=0.1)   # wave_2D.py:153-158


def lumped_matrix(prm):
    """A_em = (J - R) diag(1/L, 1/m, K)   wave_2D.py:160-166"""
    J = np.array([[0, prm['Bl_em'], 0], [-prm['Bl_em'], 0, -1], [0, 1, 0]], dtype=np.float64)
    R = np.diag([prm['R1_em'], prm['R2_em'], 0.0])
    conv = np.diag([1.0 / prm['L_em'], 1.0 / prm['m_em'], prm['K_em']])
    return (J - R).dot(conv)


def _check_case(domainShape, timeIntScheme, dirichletImp, interConnection, analytical, controlType):
    def stop(msg):
        print(msg)
        raise SystemExit
    if controlType == 'lqr':
        stop('lqr no longer implemented')
    if controlType == 'passivity' and interConnection != 'IC':
        stop('passivity control only implemented for IC interconnection')
    if controlType == 'casimir' and interConnection is not None:
        stop('casimir control not implemented for IC interconnection')
    if analytical and domainShape != 'R':
        stop('analytical solution for rectangle domain only')
    if interConnection == 'IC' and dirichletImp != 'weak':
        stop('can only do interconnection model with weak boundary conditions')
    if domainShape not in ('R', 'S_1C'):
        stop('domainShape \"{}\" hasn\'t been created'.format(domainShape))
    if interConnection == 'IC' and timeIntScheme not in ['SV', 'SE', 'SM']:
        stop('only SE, SV and SM time int schemes implemented for interconnection model')
    if timeIntScheme not in ['SV', 'IE', 'EE', 'SE', 'SM', 'EH']:
        stop('time integration scheme {} not implemented'.format(timeIntScheme))
    if dirichletImp not in ('weak', 'strong'):
        stop('dirichlet implementation {} not implemented'.format(dirichletImp))


class _Solver:
    """assemble(a) once + LU solve per step (wave_2D.py:586-590,820,854)."""

    def __init__(self, A, refactor_every_solve=False):
        self.A = sp.csc_matrix(A)
        self.refactor = refactor_every_solve
        self.lu = None if refactor_every_solve else spla.splu(self.A)

    def solve(self, b):
        if self.refactor:
            return spla.splu(self.A).solve(b)
        return self.lu.solve(b)


def wave_2D_solve_oracle(tFinal, numSteps, mesh, xLength, yLength,
                         domainShape='R', timeIntScheme='SV', dirichletImp='weak',
                         K_wave=1, rho=1, interConnection=None, analytical=False,
                         controlType=None, input_stop_t=None, control_start_t=None,
                         basis_order=(1, 1), lumped=None, refactor_every_solve=False,
                         return_state=False, timing=None):
    """Same arguments / H_array columns as the reference (wave_2D.py:15-20,1011-1032), with the
    mesh passed in.  Returns (H_array, numCells) or, with return_state, also a dict holding the
    final p, q, x (oracle numbering = fem.Topology numbering)."""
    _check_case(domainShape, timeIntScheme, dirichletImp, interConnection, analytical, controlType)
    vertices, cells = mesh
    topo = fem.Topology(vertices, cells)
    tags = fem.tag_boundary(topo, domainShape, xLength, yLength)
    nv, ne = topo.n_vertices, topo.n_edges
    high_order = tuple(basis_order) != (1, 1)
    if high_order:
        # basis_order of wave_2D.py:15-20,174-175: P_k x RT_k spaces, same forms; nv / ne below are then the numbers of
        # P / RT dofs.  Restated for the weak form without control, every scheme, with or without the lumped model (the reference
        # itself runs the explicit analytical variants at higher order:
        # main_wave_2D.py:53-114, run_test_cases.py:30-31,44-46).
        if dirichletImp != 'weak' or controlType is not None:
            raise NotImplementedError('oracle: higher-order pairs are restated for the weak form without control')
        from . import fem_ho
        spaces = fem_ho.SpacesHO(topo, int(basis_order[0]), int(basis_order[1]))
        nv, ne = spaces.n_p, spaces.n_q
    IC = interConnection == 'IC'
    strong = dirichletImp == 'strong'
    casimir = controlType == 'casimir'
    scheme = timeIntScheme

    dt = tFinal / numSteps                                   # wave_2D.py:131
    numIntSubSteps = 2 if scheme in ['EH', 'SV'] else 1      # wave_2D.py:135-137
    K = float(K_wave)
    rho_val = float(rho)
    c2 = K / rho_val
    c = np.sqrt(c2)

    # ---- boundary data (wave_2D.py:372-409) ------------------------------------------------------
    if analytical:
        omega = c * np.sqrt(np.pi ** 2 / (4 * xLength ** 2) + 4 * np.pi ** 2 / yLength ** 2)   # cSqrtConst :375
        left_weight = lambda x, y: rho_val * omega * np.sin(2 * np.pi * y / yLength + np.pi / 2)
        g_time = lambda t: np.cos(t * omega)
    else:
        left_weight = None
        if casimir:
            g_time = lambda t: 3 * np.sin(8 * np.pi * t)                                      # :380
        else:
            g_time = lambda t: 10 * np.sin(8 * np.pi * t) if t < 0.25 else 0.0               # :384
    if high_order:
        ops = spaces.assemble(tags, left_weight=left_weight)
    else:
        ops = fem.assemble_wave_operators(topo, tags, left_weight=left_weight)
    Mp, Mq, bL = ops['M_p'], ops['M_q'], ops['b_L']
    if strong:
        # (phi, grad psi) = -(div phi, psi) + <phi.n, psi>_whole boundary; keep the quadrature-assembled G
        G = ops['G']
    D = ops['D']

    if IC:
        prm = dict(DEFAULT_LUMPED)
        if lumped:
            prm.update(lumped)
        A_em = lumped_matrix(prm)
        m_em, L_em, K_em = prm['m_em'], prm['L_em'], prm['K_em']
        if controlType == 'passivity':
            def u_input(t, h_1):                                                             # control_input.py:12-18
                if t < input_stop_t:
                    return 10 * np.sin(8 * np.pi * t)
                elif t < control_start_t:
                    return 0.0
                return -100.0 * h_1
        else:
            def u_input(t, h_1):                                                             # :433,444
                return 5 * np.sin(8 * np.pi * t) if t < 0.25 else 0.0

    nx3 = 3 if IC else 0
    ntot = nv + ne + nx3
    sl_p, sl_q, sl_x = slice(0, nv), slice(nv, nv + ne), slice(nv + ne, ntot)

    # strong Dirichlet rows (wave_2D.py:413-422): P-dofs on the left (value g) and right (value 0) sides
    if strong:
        be = topo.boundary_edges
        left_v = np.unique(topo.edges[be[tags == fem.TAG_LEFT]].ravel())
        right_v = np.unique(topo.edges[be[tags == fem.TAG_RIGHT]].ravel())
        if analytical:
            left_vals_w = left_weight(topo.vertices[left_v, 0], topo.vertices[left_v, 1])
        else:
            left_vals_w = np.ones(left_v.shape[0])
        bc_rows = np.concatenate([left_v, right_v])

    def apply_bc_matrix(A):
        if not strong:
            return A
        A = sp.lil_matrix(A)
        for r in bc_rows:
            A.rows[r] = [r]
            A.data[r] = [1.0]
        return A.tocsr()

    def apply_bc_rhs(b, t):
        if strong:
            b[left_v] = left_vals_w * g_time(t)
            b[right_v] = 0.0
        return b

    # ---- block assembly of lhs(F) (wave_2D.py:449-595) ---------------------------------------------
    def assemble_system(dt_p, dt_q, ap, aq, lumped_rows):
        """lhs of:  M_p(p - p_m) + dt_p K [D^T q1 (+ casimir)] = 0 ;  M_q(q - q_old) - dt_q/rho [D p1 - b_L pLeft (- casimir)] = 0
        with p1 = ap p + (1-ap) p_m, q1 = aq q + (1-aq) q_m (weak) or the G-variants (strong);
        lumped_rows: None or dict(Axx [3,3] lhs block, q_to_x1 coefficient on b_L^T q in row 1, x1_to_q coefficient on b_L x1)."""
        App = Mp.copy()
        Aqq = Mq.copy()
        if strong:
            Apq = -dt_p * K * aq * G.T
            Aqp = dt_q / rho_val * ap * G
        else:
            Apq = dt_p * K * aq * D.T
            Aqp = -dt_q / rho_val * ap * D
        if casimir:
            App = App + dt_p * K * 0.5 * ops['B_top']
            Aqq = Aqq + dt_q / rho_val * 0.5 * ops['B_right']
        blocks = [[App, Apq], [Aqp, Aqq]]
        if lumped_rows is not None:
            col_x1 = sp.csr_matrix((bL * lumped_rows['x1_to_q'], (np.arange(ne), np.ones(ne, dtype=int))), shape=(ne, 3))
            row_x1 = sp.csr_matrix((bL * lumped_rows['q_to_x1'], (np.ones(ne, dtype=int), np.arange(ne))), shape=(3, ne))
            blocks = [[App, Apq, None], [Aqp, Aqq, col_x1], [None, row_x1, sp.csr_matrix(lumped_rows['Axx'])]]
        A = sp.bmat(blocks, format='csr')
        return _Solver(apply_bc_matrix(A), refactor_every_solve)

    I3 = np.eye(3)
    if numIntSubSteps == 1:
        ap, aq = _ALPHA[scheme]
        lr = None
        if IC:
            if scheme == 'SE':          # wave_2D.py:552-556 : new x0,x1 implicit, old x2; pLeft = rho x1/m (:397)
                Aimp = A_em.copy()
                Aimp[:, 2] = 0.0
                lr = dict(Axx=I3 - dt * Aimp, q_to_x1=-dt * K * aq, x1_to_q=dt / m_em)
            else:                       # SM  wave_2D.py:562-572 ; pLeft = 0.5 rho (x1_m + x1)/m (:400)
                lr = dict(Axx=I3 - 0.5 * dt * A_em, q_to_x1=-dt * K * aq, x1_to_q=0.5 * dt / m_em)
        sysA = assemble_system(dt, dt, ap, aq, lr)
    elif scheme == 'SV':
        # stage 1 = half explicit step (F_temp, :452-455 / :463-472 / :530-534); p1,q1 old => no coupling blocks
        lr1 = None
        lr2 = None
        if IC:
            A1 = np.zeros((3, 3))
            A1[:, 2] = A_em[:, 2]                                  # only x[2] is implicit in F_temp_em
            lr1 = dict(Axx=I3 - 0.5 * dt * A1, q_to_x1=0.0, x1_to_q=0.0)
            A2 = A_em.copy()
            A2[:, 2] = 0.0                                          # F_em :538-546
            A2[2, :] = A_em[2, :] * 1.0
            A2[2, 2] = 0.0
            lr2 = dict(Axx=I3 - 0.5 * dt * A2, q_to_x1=0.0, x1_to_q=0.5 * dt / m_em)
        sysT = assemble_system(0.5 * dt, 0.5 * dt, 0.0, 0.0, lr1)
        # stage 2 (F, :457-460 / :474-483): q1 = q_temp (known), p1 = p (new) on a half step
        sysA = assemble_system(dt, 0.5 * dt, 1.0, 0.0, lr2)
    else:  # EH: two explicit systems with identical lhs (:487-518)
        sysT = assemble_system(dt, dt, 0.0, 0.0, None)
        sysA = sysT

    # ---- rhs(F) ------------------------------------------------------------------------------------
    def rhs_wave(dt_p, dt_q, p_old, q_old, p_star_old, q_star_old, pleft_known, t_bc, q_base=None):
        """explicit part: M_p p_old - dt_p K D^T q*_old ;  M_q q_base + dt_q/rho (D p*_old - b_L pleft_known)."""
        b = np.zeros(ntot)
        if q_base is None:
            q_base = q_old
        if strong:
            b[sl_p] = Mp @ p_old + dt_p * K * (G.T @ q_star_old)
            b[sl_q] = Mq @ q_base - dt_q / rho_val * (G @ p_star_old)
        else:
            b[sl_p] = Mp @ p_old - dt_p * K * (D.T @ q_star_old)
            b[sl_q] = Mq @ q_base + dt_q / rho_val * (D @ p_star_old - bL * pleft_known)
        return b

    # ---- initial conditions (wave_2D.py:732-771) -----------------------------------------------------
    p_m = np.zeros(nv)
    q_m = np.zeros(ne)
    x_m = np.zeros(3)
    H_init = 0.0
    X, Y = (spaces.p_coords[:, 0], spaces.p_coords[:, 1]) if high_order else (topo.vertices[:, 0], topo.vertices[:, 1])
    if analytical:
        p_space = rho_val * omega * np.sin(np.pi * X / (2 * xLength) + np.pi / 2) * np.sin(2 * np.pi * Y / yLength + np.pi / 2)
        p_m = p_space.copy()
        H_init = 0.5 * p_m @ (Mp @ p_m) / rho_val                                        # :746
        if high_order:
            diag = fem_ho.AnalyticalDiagnosticsHO(spaces, rho_val, omega, xLength, yLength)
        else:
            diag = _AnalyticalDiagnostics(topo, rho_val, omega, xLength, yLength)
    elif controlType == 'passivity':
        p_m = np.sin(4 * np.pi * X / xLength)                                            # :759-763
        mid = 0.5 * (topo.vertices[topo.edges[:, 0]] + topo.vertices[topo.edges[:, 1]])
        tang = topo.vertices[topo.edges[:, 1]] - topo.vertices[topo.edges[:, 0]]
        nrm_scaled = np.stack([tang[:, 1], -tang[:, 0]], axis=1)                         # |e| n_e
        q_m = np.sin(4 * np.pi * mid[:, 0] / xLength) * nrm_scaled[:, 0]                 # :761-764 (FIAT RT1 dof = |e| q(mid).n)
        H_init = (0.5 * p_m @ (Mp @ p_m) + 0.5 * c2 * q_m @ (Mq @ q_m)) * rho_val        # :769

    # ---- solve loop (wave_2D.py:773-982) ---------------------------------------------------------------
    ncol = 12 if analytical else 8
    H_array = np.zeros((numSteps, ncol))
    H_em = 0.0
    disp = 0.0
    boundaryEnergy = 0.0
    inputEnergy = 0.0
    t = 0.0
    u_prev_x0 = 0.0     # out_xode_0 of the previous step (:808)
    tic = time.perf_counter()
    for n in range(numSteps):
        if numIntSubSteps == 1:
            t += dt                                                                      # :790-791
        elif scheme != 'SV':
            t += dt                                                                      # :797-799 (EH)
        h_1 = 0.0
        if IC and controlType == 'passivity' and t >= control_start_t:
            h_1 = u_prev_x0                                                              # :805-808

        if numIntSubSteps > 1:
            # ---------- first sub-step (:812-820) ----------
            if scheme == 'SV':
                if IC:
                    pleft0 = rho_val * x_m[1] / m_em                                     # :392
                    b = rhs_wave(0.5 * dt, 0.5 * dt, p_m, q_m, p_m, q_m, pleft0, t)
                    b[sl_x] = x_m + 0.5 * dt * (A_em[:, 0] * x_m[0] + A_em[:, 1] * x_m[1]
                                                + np.array([u_input(t, h_1), K * (bL @ q_m), 0.0]))
                else:
                    b = rhs_wave(0.5 * dt, 0.5 * dt, p_m, q_m, p_m, q_m, g_time(t), t)
            else:   # EH
                b = rhs_wave(dt, dt, p_m, q_m, p_m, q_m, g_time(t), t)
            u_temp = sysT.solve(apply_bc_rhs(b, t))
            p_temp, q_temp = u_temp[sl_p], u_temp[sl_q]
            x_temp = u_temp[sl_x]
            if not IC:
                # :830  0.5 dt c^2 (q0_.n) pLeft0_ ds(0), q0_ = q_temp (SV) / q_m (EH)
                q0_ = q_temp if scheme == 'SV' else q_m
                boundaryEnergy += 0.5 * dt * c2 * (bL @ q0_) * g_time(t)
            if scheme == 'SV':
                t += dt                                                                  # :833-834
            # ---------- second sub-step (:849-854) ----------
            if scheme == 'SV':
                if IC:
                    b = rhs_wave(dt, 0.5 * dt, p_m, q_m, p_m, q_temp, 0.0, t, q_base=q_temp)
                    u_now = u_input(t, h_1)
                    xr = np.zeros(3)
                    xr[0] = x_m[0] + dt * (0.5 * (A_em[0, 0] * x_m[0] + A_em[0, 1] * x_m[1]) + A_em[0, 2] * x_temp[2] + u_now)
                    xr[1] = x_m[1] + dt * (0.5 * (A_em[1, 0] * x_m[0] + A_em[1, 1] * x_m[1]) + A_em[1, 2] * x_temp[2]
                                           + K * (bL @ q_temp))
                    xr[2] = x_temp[2] + dt * 0.5 * A_em[2, 2] * x_temp[2]
                    b[sl_x] = xr
                else:
                    b = rhs_wave(dt, 0.5 * dt, p_m, q_m, p_m, q_temp, g_time(t), t, q_base=q_temp)
            else:   # EH: q1 = 0.5 (q_m + q_temp), p1 = 0.5 (p_m + p_temp)
                b = rhs_wave(dt, dt, p_m, q_m, 0.5 * (p_m + p_temp), 0.5 * (q_m + q_temp), g_time(t), t)
            # SV stage 2 has p1 = p implicit in the q rows: move it to the lhs (done in sysA), so zero it here
            if scheme == 'SV':
                if strong:
                    b[sl_q] = Mq @ q_temp
                else:
                    pl = 0.0 if IC else g_time(t)
                    b[sl_q] = Mq @ q_temp - 0.5 * dt / rho_val * (bL * pl)
            u_new = sysA.solve(apply_bc_rhs(b, t))
        else:
            ap, aq = _ALPHA[scheme]
            if IC:
                u_now = u_input(t, h_1)
                if scheme == 'SE':
                    b = rhs_wave(dt, dt, p_m, q_m, (1 - ap) * p_m, (1 - aq) * q_m, 0.0, t)
                    b[sl_x] = x_m + dt * (A_em[:, 2] * x_m[2] + np.array([u_now, K * (1 - aq) * (bL @ q_m), 0.0]))
                else:   # SM
                    b = rhs_wave(dt, dt, p_m, q_m, (1 - ap) * p_m, (1 - aq) * q_m, 0.5 * rho_val * x_m[1] / m_em, t)
                    b[sl_x] = x_m + dt * (0.5 * A_em @ x_m + np.array([u_now, K * (1 - aq) * (bL @ q_m), 0.0]))
            else:
                b = rhs_wave(dt, dt, p_m, q_m, (1 - ap) * p_m, (1 - aq) * q_m, g_time(t), t)
                if casimir:
                    b[sl_p] -= dt * K * 0.5 * (ops['B_top'] @ p_m)
                    b[sl_q] -= dt / rho_val * 0.5 * (ops['B_right'] @ q_m)
            u_new = sysA.solve(apply_bc_rhs(b, t))

        p_, q_ = u_new[sl_p], u_new[sl_q]
        x_ = u_new[sl_x] if IC else np.zeros(3)

        # ---------- energies (:876-927) ----------
        H_wave = 0.5 * p_ @ (Mp @ p_) / rho_val + 0.5 * K * q_ @ (Mq @ q_)              # :882,907
        if numIntSubSteps == 1:
            q1_ = {'EE': q_m, 'IE': q_, 'SE': q_m, 'SM': 0.5 * (q_m + q_)}[scheme]     # :603-615
        else:
            q1_ = q_temp                                                                 # :602,618
        if not IC:
            if numIntSubSteps == 1:
                bE = dt * c2 * (bL @ q1_) * g_time(t)                                    # :631
                if casimir:
                    p1_ = 0.5 * (p_m + p_)
                    bE += dt * c2 * (p1_ @ (ops['B_top'] @ p1_))                         # :633
                    bE += dt * c2 * (q1_ @ (ops['B_right'] @ q1_))                       # :634
                boundaryEnergy += bE
            else:
                boundaryEnergy += 0.5 * dt * c2 * (bL @ q1_) * g_time(t)                 # :625,880
            E = H_wave + boundaryEnergy - H_init
            H = H_wave
        else:
            u_now = u_input(t, h_1)
            if scheme == 'SV':
                inp = 0.5 * dt * u_now * x_m[0] / L_em + 0.5 * dt * u_now * x_[0] / L_em   # :828,892 (both assembled at :903)
            elif scheme == 'SE':
                inp = dt * u_now * x_[0] / L_em                                          # :894
            else:   # SM
                inp = 0.5 * dt * u_now * (x_m[0] + x_[0]) / L_em                         # :898
            inputEnergy += inp                                                           # :903
            H_em = 0.5 * (x_[0] ** 2 / L_em + x_[1] ** 2 / m_em + K_em * x_[2] ** 2)     # :911-912
            E = -inputEnergy + H_wave + H_em - H_init                                    # :914
            H = H_wave + H_em
            if scheme == 'SV':                                                           # :622-628 with :392-395
                boundaryEnergy += 0.5 * dt * c2 * (bL @ q_temp) * rho_val * (x_m[1] + x_[1]) / m_em
            elif scheme == 'SE':                                                         # :631 with :398
                boundaryEnergy += dt * c2 * (bL @ q1_) * rho_val * x_[1] / m_em
            else:                                                                        # :631 with :401
                boundaryEnergy += dt * c2 * (bL @ q1_) * 0.5 * rho_val * (x_m[1] + x_[1]) / m_em
            disp = x_[2]                                                                 # :921
            u_prev_x0 = x_[0]

        row = [t, H, E, disp, boundaryEnergy, inputEnergy, H_em, H_wave]
        if analytical:
            row += list(diag.evaluate(p_, t))
        H_array[n, :] = row
        p_m, q_m, x_m = p_, q_, x_                                                       # :982
    if timing is not None:
        timing['loop_seconds'] = time.perf_counter() - tic
        timing['n_dof'] = ntot
    if return_state:
        return H_array, topo.n_cells, dict(p=p_m, q=q_m, x=x_m, topo=topo, tags=tags, ops=ops,
                                           spaces=spaces if high_order else None)
    return H_array, topo.n_cells


class _AnalyticalDiagnostics:
    """wave_2D.py:929-965: errornorm (P_{k+3} interpolation of both fields), nodal max error, point probe."""

    def __init__(self, topo, rho_val, omega, xLength, yLength, k=1, xPoint=0.155, yPoint=0.189):
        self.topo = topo
        self.space = lambda x, y: rho_val * omega * np.sin(np.pi * x / (2 * xLength) + np.pi / 2) * \
            np.sin(2 * np.pi * y / yLength + np.pi / 2)
        self.omega = omega
        kk = k + 3
        nodes = fem.lagrange_nodes(kk)                                        # reference (x,y)
        lam = np.stack([1 - nodes[:, 0] - nodes[:, 1], nodes[:, 0], nodes[:, 1]], axis=1)
        V = topo.vertices[topo.cells]
        xy = np.einsum('nj,cjd->cnd', lam, V)                                 # physical P4 nodes per cell
        self.exact_nodes = self.space(xy[..., 0], xy[..., 1])                 # [nc, 15]
        self.lam_nodes = lam
        qx, qy, qw = fem.high_order_tri_rule(6)                               # exact to degree 10 >= 2*4
        tab = fem.lagrange_tabulate(kk, qx, qy)                               # [nq, 15]
        self.mass_ref = np.einsum('qi,qj,q->ij', tab, tab, qw)                # reference P4 mass (|K| = 1)
        self.exact_vertices = self.space(topo.vertices[:, 0], topo.vertices[:, 1])
        # point probe: locate the containing cell once
        P = np.array([xPoint, yPoint])
        T = np.stack([V[:, 1] - V[:, 0], V[:, 2] - V[:, 0]], axis=2)          # [nc,2,2]
        rhs = P[None, :] - V[:, 0]
        det = T[:, 0, 0] * T[:, 1, 1] - T[:, 0, 1] * T[:, 1, 0]
        l1 = (rhs[:, 0] * T[:, 1, 1] - rhs[:, 1] * T[:, 0, 1]) / det
        l2 = (T[:, 0, 0] * rhs[:, 1] - T[:, 1, 0] * rhs[:, 0]) / det
        l0 = 1 - l1 - l2
        inside = np.nonzero((l0 >= -1e-12) & (l1 >= -1e-12) & (l2 >= -1e-12))[0]
        self.pt_cell = int(inside[0]) if inside.size else -1
        if self.pt_cell >= 0:
            self.pt_lam = np.array([l0[self.pt_cell], l1[self.pt_cell], l2[self.pt_cell]])

    def evaluate(self, p_h, t):
        ct = np.cos(t * self.omega)
        ph_nodes = p_h[self.topo.cells] @ self.lam_nodes.T                    # P1 field at P4 nodes [nc,15]
        e = self.exact_nodes * ct - ph_nodes
        err_L2 = np.sqrt(np.sum(np.einsum('ci,ij,cj->c', e, self.mass_ref, e) * self.topo.area))
        err_max = np.abs(self.exact_vertices * ct - p_h).max()                # :958
        if self.pt_cell >= 0:
            vc = self.topo.cells[self.pt_cell]
            p_point = float(p_h[vc] @ self.pt_lam)                            # :949
            p_point_exact = float((self.exact_vertices[vc] * ct) @ self.pt_lam)   # :950
        else:
            p_point = p_point_exact = 0.0
        return err_L2, err_max, p_point, p_point_exact
