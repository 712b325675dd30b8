"""Drop-in for the reference solver call: ``wave_2D_solve`` (reference
``fenics_projects/wave_eqn/wave_2D.py:15-20,1032``) with the same positional/keyword arguments, the
same printed messages on invalid case combinations (``print`` + ``quit()``, :63-83,148-151) and the
same ``(H_array, numCells)`` return value (:1011-1032).  Behind it there is no dolfin/PETSc: the mesh
goes to libphwave.so (C ABI, include/phwave.h) and the whole time loop runs on the GPU.

Extra keyword arguments (defaults keep the reference behaviour): ``mesh=(vertices, cells)`` to bypass
the built-in generator (the reference calls mshr, :93-119), ``mesh_kind``, ``device``, ``patch_cells``,
``tol``, ``lumped`` (dict overriding the hard-coded constants of :153-158), ``solver_flags`` (phw_params.flags; default:
pipelined kernels for meshes of >= 200 000 cells, the single-launch whole-loop kernel below), ``return_state``, ``save_every``
(with ``saveP``: write p every that many steps to ``<outputDir>/p.xdmf``; default ``None`` = no field output, because a
per-step device-to-host copy would dominate the run).
"""
import os
import time

import numpy as np

from . import _lib
from .mesh import generate_mesh
from .solver import Mesh, Solver, tag_boundary_edges, port_weights

_DEFAULT_LUMPED = dict(Bl_em=5.0, R1_em=0.0, R2_em=0.0, K_em=5.0, m_em=0.15, L_em=0.1)   # wave_2D.py:153-158


def _quit(msg):
    print(msg)
    raise SystemExit


def wave_2D_solve(tFinal, numSteps, outputDir,
                  nx, xLength, yLength,
                  domainShape='R', timeIntScheme='SV', dirichletImp='weak',
                  K_wave=1, rho=1, BOUNDARYPLOT=False, interConnection=None,
                  analytical=False, saveP=True, controlType=None,
                  input_stop_t=None, control_start_t=None, basis_order=(1, 1),
                  mesh=None, mesh_kind='structured', device=0, patch_cells=None, tol=1e-13, extrap_order=3, save_every=None,
                  lumped=None, return_state=False, verbose=True, solver_flags=None):
    """Solve the 2D port-Hamiltonian wave equation; see the reference docstring (wave_2D.py:22-51).

    Returns (H_array [numSteps, 8 | 12], numCells)."""
    rank = 0
    # ---- check case combinations (wave_2D.py:63-83) ----
    if controlType == 'lqr':
        _quit('lqr no longer implemented')
    if controlType == 'passivity':
        if interConnection != 'IC':
            _quit('passivity control only implemented for IC interconnection')
    if controlType == 'casimir':
        if interConnection is not None:
            _quit('casimir control not implemented for IC interconnection')
    if analytical:
        if not domainShape == 'R':
            _quit('analytical solution for rectangle domain only')
    if interConnection == 'IC':
        if not dirichletImp == 'weak':
            _quit('can only do interconnection model with weak boundary conditions')
    BOUNDARYPLOT = False                                               # wave_2D.py:86
    tic = time.time()
    if rank == 0 and verbose:
        print('Started solving for case: {}, {}, {} '.format(domainShape, timeIntScheme, dirichletImp))
    if domainShape not in ('R', 'S_1C'):
        _quit('domainShape \"{}\" hasn\'t been created'.format(domainShape))
    if interConnection == 'IC' and timeIntScheme not in ['SV', 'SE', 'SM']:
        _quit('only SE, SV and SM time int schemes implemented for interconnection model')
    if timeIntScheme not in _lib.SCHEMES:
        _quit('time integration scheme {} not implemented'.format(timeIntScheme))
    if dirichletImp not in ('weak', 'strong'):
        _quit('dirichlet implementation {} not implemented'.format(dirichletImp))
    if tuple(basis_order) != (1, 1):
        return _solve_high_order(tFinal, numSteps, outputDir, nx, xLength, yLength, domainShape, timeIntScheme, dirichletImp,
                                 K_wave, rho, interConnection, analytical, controlType, tuple(basis_order), mesh, mesh_kind,
                                 device, tol, extrap_order, lumped, return_state, verbose, tic, solver_flags or 0, saveP=saveP, save_every=save_every)
    if controlType == 'casimir' and timeIntScheme != 'SM':
        raise NotImplementedError('casimir control is only defined for the SM scheme (wave_2D.py:615)')
    if controlType == 'passivity' and timeIntScheme in ['SV', 'EH']:
        raise NotImplementedError('passivity control is only wired for one-sub-step schemes (wave_2D.py:844)')

    # ---- mesh (wave_2D.py:92-128) ----
    if mesh is None:
        mesh = generate_mesh(domainShape, nx, xLength, yLength, kind=mesh_kind)
    vertices, cells = mesh
    # meshes far beyond the caches run the bulk-copy pipelined kernels (csrc/kernels_pipe.cuh; flags bits 8, 9, 11, 12), which need
    # patches of <= ~1000 cells so that two shared-memory stages fit; small (cache-resident) meshes keep the lean kernels
    big = len(cells) >= 200000
    if solver_flags is None:
        # (the explicit schemes are what the pipelined kernels have been measured and parity-tested with; the implicit block
        # solve has no pipelined variant yet)
        explicit = timeIntScheme in ('EE', 'SE', 'EH', 'SV') and controlType != 'casimir'
        solver_flags = (_lib.FLAGS_PIPELINED if big and explicit and (patch_cells is None or patch_cells <= 1024) else 0)
    # cache-resident meshes (everything the reference itself runs): the whole time loop in ONE launch (csrc/kernels_loop.cuh; flags
    # bit 15), the mesh cut into <= 16 patches = one thread-block cluster; PHW_NO_LOOP=1 keeps the multi-launch path (A/B, tests)
    # (measured, profiles/r2g_*: the coupled block solve of IE / SM runs its row passes from L2 and wants every SM once the mesh has
    # more than ~12 000 cells - there the multi-launch path is faster than one 16-CTA cluster)
    block_big = timeIntScheme in ('IE', 'SM') and len(cells) > 12000
    if solver_flags == 0 and not big and not analytical and patch_cells is None and not block_big and not os.environ.get('PHW_NO_LOOP'):
        solver_flags = _lib.FLAGS_LOOP
        patch_cells = _lib.loop_patch_cells(len(cells)) or 2048
    if patch_cells is None:
        patch_cells = (1024 if solver_flags & 256 else 2048) if big else max(64, min(2048, len(cells) // 256))
    m = Mesh(vertices, cells, patch_cells=patch_cells)
    numCells = m.n_cells
    if rank == 0 and verbose:
        print('number of cells = {}'.format(numCells))

    dt_value = tFinal / numSteps                                       # wave_2D.py:131
    c = np.sqrt(K_wave / rho)
    prm = _lib.PhwParams()
    prm.scheme = _lib.SCHEMES[timeIntScheme]
    prm.strong = 1 if dirichletImp == 'strong' else 0
    prm.interconnection = 1 if interConnection == 'IC' else 0
    prm.casimir = 1 if controlType == 'casimir' else 0
    prm.analytical = 1 if analytical else 0
    prm.dt, prm.K_wave, prm.rho = dt_value, float(K_wave), float(rho)
    prm.tol, prm.max_iter, prm.flags, prm.extrap_order = float(tol), 2000, int(solver_flags), int(extrap_order)   # cap only: see DESIGN.md section 2 (stretched cells)
    lp = dict(_DEFAULT_LUMPED)
    if lumped:
        lp.update(lumped)
    for k, v in lp.items():
        setattr(prm, k, float(v))

    # ---- boundary markers and data (wave_2D.py:226-285,326-444) ----
    tags = tag_boundary_edges(m, domainShape, xLength, yLength)
    weights = None
    omega = 0.0
    if analytical:
        omega = c * np.sqrt(np.pi ** 2 / (4 * xLength ** 2) + 4 * np.pi ** 2 / yLength ** 2)           # cSqrtConst :375
        weights = port_weights(m, tags, lambda x, y: rho * omega * np.sin(2 * np.pi * y / yLength + np.pi / 2))
        prm.input_kind, prm.in_amp, prm.in_omega = _lib.IN_COSINE, 1.0, omega                         # :376
    elif interConnection == 'IC':
        if controlType == 'passivity':
            prm.input_kind = _lib.IN_PASSIVITY                                                         # control_input.py:12-18
            prm.in_amp, prm.in_omega, prm.in_gain = 10.0, 8 * np.pi, 100.0
            prm.in_t_stop, prm.in_t_ctrl = float(input_stop_t), float(control_start_t)
        else:
            prm.input_kind, prm.in_amp, prm.in_omega, prm.in_t_stop = _lib.IN_SINE_WINDOW, 5.0, 8 * np.pi, 0.25   # :433,444
    elif controlType == 'casimir':
        prm.input_kind, prm.in_amp, prm.in_omega = _lib.IN_SINE, 3.0, 8 * np.pi                        # :380
    else:
        prm.input_kind, prm.in_amp, prm.in_omega, prm.in_t_stop = _lib.IN_SINE_WINDOW, 10.0, 8 * np.pi, 0.25    # :384
    m.set_boundary(tags, weights)
    if dirichletImp == 'strong':
        # DirichletBC(U.sub(0), pLeftBoundary1, forced_boundary) / DirichletBC(U.sub(0), 0, fixed_boundary)  (:413-420)
        ids, _ = m.boundary()
        edges = m.edges()[0]
        left_v = np.unique(edges[ids[tags == 0]].ravel())
        right_v = np.unique(edges[ids[tags == 2]].ravel())
        if analytical:
            left_w = rho * omega * np.sin(2 * np.pi * m.vertices[left_v, 1] / yLength + np.pi / 2)
        else:
            left_w = np.ones(left_v.shape[0])
        m.set_dirichlet(np.concatenate([left_v, right_v]), np.concatenate([left_w, np.zeros(right_v.shape[0])]))

    solver = Solver(m, prm, device=device)
    # mesh quality: the mass solves are iterative here (the reference's LU is not), and their cost follows the worst cell
    st0 = solver.stats()
    if st0['lam_max_q'] > 10.0 * st0['lam_min_q'] and rank == 0:
        import warnings
        warnings.warn('stretched cells in the mesh: the element-wise bound of the condition number of the Jacobi-scaled RT1 mass matrix is '
                      '{:.0f} (quality meshes: <= 4); the M_q solves switch from Chebyshev iteration to conjugate gradients where available '
                      '(one GPU, weak form) and need on the order of sqrt(kappa) x 15 iterations per solve'
                      .format(st0['lam_max_q'] / st0['lam_min_q']), UserWarning, stacklevel=2)

    # ---- initial conditions (wave_2D.py:732-771) ----
    H_init = 0.0
    X, Y = m.vertices[:, 0], m.vertices[:, 1]
    p_space = None
    if analytical:
        p_space = rho * omega * np.sin(np.pi * X / (2 * xLength) + np.pi / 2) * np.sin(2 * np.pi * Y / yLength + np.pi / 2)
        solver.set_state(p=p_space, q=np.zeros(m.n_edges), x3=np.zeros(3))
        Hp, Hq = solver.energy_parts()
        H_init = Hp                                                     # :746 (q_m = 0)
    elif controlType == 'passivity':
        edges = m.edges()[0]
        p0 = np.sin(4 * np.pi * X / xLength)                            # :759-763
        mid = 0.5 * (m.vertices[edges[:, 0]] + m.vertices[edges[:, 1]])
        tang = m.vertices[edges[:, 1]] - m.vertices[edges[:, 0]]
        q0 = np.sin(4 * np.pi * mid[:, 0] / xLength) * tang[:, 1]       # RT1 dof = |e| q(mid).n_e, n_e = (t_y, -t_x)/|e|
        solver.set_state(p=p0, q=q0, x3=np.zeros(3))
        Hp, Hq = solver.energy_parts()
        # :769  rho (1/2 p^T M_p p + 1/2 c^2 q^T M_q q)   (deliberately not the H_wave formula of :882)
        H_init = rho * (rho * Hp) + rho * (Hq / K_wave) * (K_wave / rho)
    solver.set_H_init(H_init)

    # ---- time loop on the device (wave_2D.py:787-982) ----
    if analytical:
        diag = _AnalyticalDiagnostics(m, p_space, omega)
        pt_v = m.cells[diag.pt_cell].astype(np.int32) if diag.pt_cell >= 0 else -np.ones(3, dtype=np.int32)
        pt_l = diag.pt_lam if diag.pt_cell >= 0 else np.zeros(3)
        mass = np.ascontiguousarray(diag.mass_ref, dtype=np.float64)
        lam = np.ascontiguousarray(diag.lam, dtype=np.float64)
        ev = np.ascontiguousarray(p_space, dtype=np.float64)
        pt_v = np.ascontiguousarray(pt_v, dtype=np.int32)
        pt_l = np.ascontiguousarray(pt_l, dtype=np.float64)
        _lib.check(solver.lib.phw_set_analytical(solver.handle, float(rho * omega), float(np.pi / (2 * xLength)), float(2 * np.pi / yLength),
                                                 float(omega), mass.ctypes.data_as(_lib.c_f64p), lam.ctypes.data_as(_lib.c_f64p),
                                                 ev.ctypes.data_as(_lib.c_f64p), pt_v.ctypes.data_as(_lib.c_i32p), pt_l.ctypes.data_as(_lib.c_f64p)))
    if saveP and save_every and outputDir:
        H_array = _run_saving_fields(solver, numSteps, int(save_every), outputDir, m.vertices, m.cells, m.n_vertices,
                                     (lambda t: p_space * np.cos(omega * t)) if analytical else None)
    else:
        H_array = solver.run(numSteps)
    stats = solver.stats()
    totalTime = time.time() - tic
    if rank == 0 and verbose:
        print('{}, {}, {} Simulation finished in {} seconds'.format(domainShape, timeIntScheme, dirichletImp, totalTime))
    if return_state:
        p, q, x = solver.get_state()
        solver.close()
        return H_array, numCells, dict(p=p, q=q, x=x, stats=stats, mesh=m)
    solver.close()
    m.close()
    return H_array, numCells


def _run_saving_fields(solver, numSteps, save_every, outputDir, vertices, cells, n_vertices, exact_of_t):
    """The time loop with field output (wave_2D.py:646-651,865-866,964-965): p.xdmf receives the nodal p field after every
    ``save_every`` steps (the reference: every step) and, in the analytical case, p_exact.xdmf the nodal interpolant of the exact
    solution at the same instants.  The snapshots leave the device through a ring of device slots and a second stream while
    the loop goes on (phw_set_snapshots); h5py is not available, so the heavy data goes to raw little-endian binaries
    referenced from the XDMF files (ParaView reads Format="Binary")."""
    from .field_output import XdmfBinaryWriter
    H_array, snaps = solver.run_with_snapshots(numSteps, save_every)
    writer = XdmfBinaryWriter(outputDir, 'p', vertices, cells)
    writer_exact = XdmfBinaryWriter(outputDir, 'p_exact', vertices, cells) if exact_of_t is not None else None
    for k in range(snaps.shape[0]):
        t = H_array[(k + 1) * save_every - 1, 0]
        writer.write(snaps[k, :n_vertices], t, flush=False)
        if writer_exact is not None:
            writer_exact.write(exact_of_t(t), t, flush=False)
    writer.close()
    if writer_exact is not None:
        writer_exact.close()
    return H_array


class _AnalyticalDiagnostics:
    """One-time host tabulations for columns 8-11 of the analytical cases (wave_2D.py:929-965): the reference P4
    mass matrix and node coordinates that dolfin ``errornorm`` (degree_rise=3) implies, and the cell containing the
    probe point (0.155, 0.189).  The per-step evaluation runs on the device (diag_cells_k / diag_finish_k);
    ``evaluate`` is kept as a host cross-check."""

    def __init__(self, mesh, p_space_vertices, omega, xPoint=0.155, yPoint=0.189):
        self.omega = omega
        self.cells = mesh.cells.astype(np.int64)
        V = mesh.vertices[self.cells]
        # P4 equispaced nodes in barycentric coordinates
        k = 4
        nodes = np.array([(i / k, j / k) for j in range(k + 1) for i in range(k + 1 - j)])
        lam = np.stack([1 - nodes[:, 0] - nodes[:, 1], nodes[:, 0], nodes[:, 1]], axis=1)
        self.lam = lam
        xy = np.einsum('nj,cjd->cnd', lam, V)
        # exact spatial factor at the P4 nodes: same closed form as the initial condition (wave_2D.py:741)
        self.exact_vertices = p_space_vertices
        self._xy = xy
        self.area = 0.5 * np.abs((V[:, 1, 0] - V[:, 0, 0]) * (V[:, 2, 1] - V[:, 0, 1]) - (V[:, 1, 1] - V[:, 0, 1]) * (V[:, 2, 0] - V[:, 0, 0]))
        # reference P4 mass matrix by exact quadrature
        def mono(x, y):
            return np.stack([x ** a * y ** b for b in range(k + 1) for a in range(k + 1 - b)], axis=-1)
        gx, gw = np.polynomial.legendre.leggauss(6)
        u, wu = 0.5 * (gx + 1), 0.5 * gw
        U, W2 = np.meshgrid(u, u, indexing='ij')
        qx, qy = U.ravel(), (W2 * (1 - U)).ravel()
        qw = 2.0 * (np.outer(wu, wu) * (1 - U)).ravel()
        tab = np.linalg.solve(mono(nodes[:, 0], nodes[:, 1]).T, mono(qx, qy).T).T
        self.mass_ref = np.einsum('qi,qj,q->ij', tab, tab, qw)
        self.exact_nodes = None
        # point probe
        P = np.array([xPoint, yPoint])
        T = np.stack([V[:, 1] - V[:, 0], V[:, 2] - V[:, 0]], axis=2)
        rhs = P[None, :] - V[:, 0]
        det = T[:, 0, 0] * T[:, 1, 1] - T[:, 0, 1] * T[:, 1, 0]
        l1 = (rhs[:, 0] * T[:, 1, 1] - rhs[:, 1] * T[:, 0, 1]) / det
        l2 = (T[:, 0, 0] * rhs[:, 1] - T[:, 1, 0] * rhs[:, 0]) / det
        l0 = 1 - l1 - l2
        inside = np.nonzero((l0 >= -1e-12) & (l1 >= -1e-12) & (l2 >= -1e-12))[0]
        self.pt_cell = int(inside[0]) if inside.size else -1
        if self.pt_cell >= 0:
            self.pt_lam = np.array([l0[self.pt_cell], l1[self.pt_cell], l2[self.pt_cell]])

    def set_space(self, fn):
        self.exact_nodes = fn(self._xy[..., 0], self._xy[..., 1])

    def evaluate(self, p_h, t):
        if self.exact_nodes is None:
            raise RuntimeError('set_space must be called first')
        ct = np.cos(t * self.omega)
        e = self.exact_nodes * ct - p_h[self.cells] @ self.lam.T
        err_L2 = np.sqrt(np.sum(np.einsum('ci,ij,cj->c', e, self.mass_ref, e) * self.area))
        err_max = np.abs(self.exact_vertices * ct - p_h).max()
        if self.pt_cell >= 0:
            vc = self.cells[self.pt_cell]
            return err_L2, err_max, float(p_h[vc] @ self.pt_lam), float((self.exact_vertices[vc] * ct) @ self.pt_lam)
        return err_L2, err_max, 0.0, 0.0


def _solve_high_order(tFinal, numSteps, outputDir, nx, xLength, yLength, domainShape, timeIntScheme, dirichletImp, K_wave, rho,
                      interConnection, analytical, controlType, basis_order, mesh, mesh_kind, device, tol, extrap_order, lumped,
                      return_state, verbose, tic, solver_flags=0, saveP=False, save_every=None):
    """basis_order != (1,1) (wave_2D.py:174-175,202-203): matrix-free cell kernels on the reference-element form (fem_ref.py,
    csrc/kernels_ho.cuh) for the explicit schemes, host-assembled CSR operators (fem_ho.py) for the implicit ones."""
    from .fem_ho import HighOrderSpaces, LagrangeRef, tri_rule
    from .fem_ref import RefSpaces
    from .solver import HighOrderSolver, MatrixFreeHighOrderSolver, tag_space_boundary
    if dirichletImp != 'weak' or controlType is not None:
        raise NotImplementedError('higher-order pairs are available for the weak form without control (all six schemes, with or without the lumped model)')
    implicit = timeIntScheme in ('IE', 'SM')
    if mesh is None:
        mesh = generate_mesh(domainShape, nx, xLength, yLength, kind=mesh_kind)
    # Two device paths: matrix-free cell kernels on the reference-element form of the spaces (csrc/kernels_ho.cuh,
    # fem_ref.py) and host-assembled CSR operators + CSR kernels (fem_ho.py).  Measured (profiles/r2o_ho_cases_matrix_free_vs_csr.jsonl):
    # an iteration costs the same on both for P2 x RT2 but the reference-element basis needs ~25 % more Chebyshev iterations there,
    # while at order 3 the assembled rows (~2.5 KB per cell) lose by 1.4 x on the paper mesh and 2.4 x beyond the L2.  Default:
    # matrix-free when an order is 3, assembled rows below; PHW_HO_MF=1 / PHW_HO_CSR=1 force either.
    # Implicit schemes (IE, SM): coupled block solve on the assembled rows (solve_block_csr_k).
    if os.environ.get('PHW_HO_CSR'):
        matrix_free = False
    elif os.environ.get('PHW_HO_MF'):
        matrix_free = True
    else:
        matrix_free = max(basis_order) >= 3
    matrix_free = matrix_free and not implicit
    S = (RefSpaces if matrix_free else HighOrderSpaces)(mesh[0], mesh[1], basis_order[0], basis_order[1])
    if verbose:
        print('number of cells = {}'.format(S.nc))
    tag_of = tag_space_boundary(S, domainShape, xLength, yLength)       # facet markers (wave_2D.py:226-280) on the boundary edges
    c = np.sqrt(K_wave / rho)
    omega = c * np.sqrt(np.pi ** 2 / (4 * xLength ** 2) + 4 * np.pi ** 2 / yLength ** 2)
    left_w = (lambda x, y: rho * omega * np.sin(2 * np.pi * y / yLength + np.pi / 2)) if analytical else None
    if matrix_free:
        ops = S.boundary_operators(tag_of, left_weight=left_w)
        bp, bq = S.jacobi_bounds()
    else:
        ops = S.assemble(tag_of, left_weight=left_w)
        bp, bq = S.jacobi_bounds(ops)
    prm = _lib.PhwParams()
    prm.scheme = _lib.SCHEMES[timeIntScheme]
    prm.interconnection = 1 if interConnection == 'IC' else 0
    prm.analytical = 1 if analytical else 0
    prm.dt, prm.K_wave, prm.rho = tFinal / numSteps, float(K_wave), float(rho)
    prm.tol, prm.max_iter, prm.extrap_order, prm.flags = float(tol), 2000, int(extrap_order), int(solver_flags)
    lp = dict(_DEFAULT_LUMPED)
    if lumped:
        lp.update(lumped)
    for k, v in lp.items():
        setattr(prm, k, float(v))
    if analytical:
        prm.input_kind, prm.in_amp, prm.in_omega = _lib.IN_COSINE, 1.0, omega
    elif interConnection == 'IC':
        prm.input_kind, prm.in_amp, prm.in_omega, prm.in_t_stop = _lib.IN_SINE_WINDOW, 5.0, 8 * np.pi, 0.25
    else:
        prm.input_kind, prm.in_amp, prm.in_omega, prm.in_t_stop = _lib.IN_SINE_WINDOW, 10.0, 8 * np.pi, 0.25
    solver = (MatrixFreeHighOrderSolver if matrix_free else HighOrderSolver)(S, ops, prm, bp, bq, device=device)
    H_init = 0.0
    if analytical:
        space = lambda x, y: rho * omega * np.sin(np.pi * x / (2 * xLength) + np.pi / 2) * np.sin(2 * np.pi * y / yLength + np.pi / 2)
        p0 = S.interpolate_p(space)
        solver.set_state(p=p0, q=np.zeros(S.nq_dofs), x3=np.zeros(3))
        H_init = solver.energy_parts()[0]
        solver.set_H_init(H_init)
        # errornorm tabulations: P_k -> P_{k+3} (wave_2D.py:956), probe point (0.155, 0.189) (:946-950)
        kk = S.kp + 3
        hi = LagrangeRef(kk)
        p_tab = (lambda bary: S.ref.tab_p(bary[:, 1], bary[:, 2])) if matrix_free else S.lag.tabulate
        T = p_tab(hi.nodes_bary)                                           # [nn, ndp]
        bary, wq = tri_rule(8)
        tab = hi.tabulate(bary)
        mass = np.einsum('qi,qj,q->ij', tab, tab, wq)
        xy = np.einsum('nj,cjd->cnd', hi.nodes_bary, S.V)
        exact_nodes = space(xy[..., 0], xy[..., 1])
        P = np.array([0.155, 0.189])
        Tm = np.stack([S.V[:, 1] - S.V[:, 0], S.V[:, 2] - S.V[:, 0]], axis=2)
        rhs = P[None, :] - S.V[:, 0]
        det = Tm[:, 0, 0] * Tm[:, 1, 1] - Tm[:, 0, 1] * Tm[:, 1, 0]
        l1 = (rhs[:, 0] * Tm[:, 1, 1] - rhs[:, 1] * Tm[:, 0, 1]) / det
        l2 = (Tm[:, 0, 0] * rhs[:, 1] - Tm[:, 1, 0] * rhs[:, 0]) / det
        l0 = 1 - l1 - l2
        inside = np.nonzero((l0 >= -1e-12) & (l1 >= -1e-12) & (l2 >= -1e-12))[0]
        if inside.size:
            pc = int(inside[0])
            pw = p_tab(np.array([[l0[pc], l1[pc], l2[pc]]]))[0]
            pd = S.cell_dofs_p[pc].astype(np.int32)
        else:
            pw, pd = np.zeros(0), np.zeros(0, dtype=np.int32)
        cd = np.ascontiguousarray(S.cell_dofs_p, dtype=np.int32)
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (T, mass, exact_nodes, S.area, p0, pw)]
        pd = np.ascontiguousarray(pd, dtype=np.int32)
        _lib.check(solver.lib.phw_ho_set_analytical(solver.handle, float(omega), S.nc, S.ndp, hi.nodes_bary.shape[0],
                                                    cd.ctypes.data_as(_lib.c_i32p), arrs[0].ctypes.data_as(_lib.c_f64p),
                                                    arrs[1].ctypes.data_as(_lib.c_f64p), arrs[2].ctypes.data_as(_lib.c_f64p),
                                                    arrs[3].ctypes.data_as(_lib.c_f64p), arrs[4].ctypes.data_as(_lib.c_f64p),
                                                    int(pd.shape[0]), pd.ctypes.data_as(_lib.c_i32p), arrs[5].ctypes.data_as(_lib.c_f64p)))
    if saveP and save_every and outputDir:     # vertex values of the P_k field (the vertex dofs come first)
        pv_space = space(S.vertices[:, 0], S.vertices[:, 1]) if analytical else None
        H_array = _run_saving_fields(solver, numSteps, int(save_every), outputDir, S.vertices, S.cells, S.nv,
                                     (lambda t: pv_space * np.cos(omega * t)) if analytical else None)
    else:
        H_array = solver.run(numSteps)
    stats = solver.stats()
    if verbose:
        print('{}, {}, {} Simulation finished in {} seconds'.format(domainShape, timeIntScheme, dirichletImp, time.time() - tic))
    if return_state:
        p, q, x = solver.get_state()
        solver.close()
        return H_array, S.nc, dict(p=p, q=q, x=x, stats=stats, spaces=S, ops=ops)
    solver.close()
    return H_array, S.nc
