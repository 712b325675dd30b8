"""Batched parameter / control sweep (BASELINE.json configs[4]): many independent coupled wave + lumped cases on the
same mesh, advanced together by the same kernels.

The reference runs its case list one case after the other (main_wave_2D.py:144, one ``wave_2D_solve`` call each).
Here the mesh is replicated once per case as disjoint components of ONE device mesh ("groups"), so a single set of
launches per step serves every case; only scalars differ between cases (K_wave, rho and the lumped constants of
wave_2D.py:153-158), which the kernels look up per patch / per port edge.  There is no coupling and no collective
between cases; on several GPUs each rank simply takes a contiguous shard of the case list.
"""
import os

import numpy as np

from . import _lib
from .mesh import generate_mesh
from .solver import Mesh, Solver, tag_boundary_edges

LUMPED_KEYS = ('Bl_em', 'R1_em', 'R2_em', 'K_em', 'm_em', 'L_em')
DEFAULT_LUMPED = dict(Bl_em=5.0, R1_em=0.0, R2_em=0.0, K_em=5.0, m_em=0.15, L_em=0.1)


def replicate_mesh(vertices, cells, n_copies):
    nv = vertices.shape[0]
    V = np.tile(np.asarray(vertices, dtype=np.float64), (n_copies, 1))
    C = (np.asarray(cells, dtype=np.int64)[None, :, :] + (np.arange(n_copies, dtype=np.int64) * nv)[:, None, None]).reshape(-1, 3)
    group = np.repeat(np.arange(n_copies, dtype=np.int32), cells.shape[0])
    return V, C.astype(np.int32), group


def sweep_cases(n_cases, seed=1234):
    """The parameter set of SURVEY.md 8d config 5 (numpy default_rng(1234))."""
    rng = np.random.default_rng(seed)
    return [dict(K_wave=rng.uniform(1, 5), rho=rng.uniform(1, 3), R1_em=rng.uniform(0, 1), R2_em=rng.uniform(0, 1),
                 L_em=rng.uniform(0.05, 0.2), m_em=rng.uniform(0.1, 0.3), K_em=rng.uniform(2, 10), Bl_em=rng.uniform(2, 8))
            for _ in range(n_cases)]


def wave_sweep_solve(tFinal, numSteps, nx, xLength, yLength, cases, domainShape='R', timeIntScheme='SV',
                     mesh=None, device=0, patch_cells=None, tol=1e-13, extrap_order=2, return_solver=False, solver_flags=0):
    """Run len(cases) independent 'IC' cases (weak form, explicit scheme) at once.

    Returns (H_arrays [n_cases, numSteps, 8], numCells) - H_arrays[i] is what wave_2D_solve would return for case i."""
    if timeIntScheme not in ('SV', 'SE'):
        raise ValueError('the sweep runs the interconnected model: SV or SE')
    if mesh is None:
        mesh = generate_mesh(domainShape, nx, xLength, yLength)
    B = len(cases)
    auto = patch_cells is None and not solver_flags
    if auto and B >= 32 and not os.environ.get('PHW_SWEEP_GROUPED') and not os.environ.get('PHW_SWEEP_LOOP'):
        # default for real sweeps: ONE copy of the (assembled) operators, state [n_dof][n_cases] with the case index fastest
        # (csrc/kernels_sweep.cuh).  PHW_SWEEP_GROUPED=1 keeps the replicated-mesh layout below (A/B, cross-check).
        return _sweep_case_fastest(tFinal, numSteps, mesh, xLength, yLength, cases, domainShape, timeIntScheme, device, tol, extrap_order,
                                   return_solver)
    V, C, group = replicate_mesh(mesh[0], mesh[1], B)
    if auto and os.environ.get('PHW_SWEEP_LOOP'):
        # one thread-block cluster per case, every case advanced through ALL its steps by its own cluster in one launch
        # (csrc/kernels_loop.cuh, solves resident in shared memory).  Measured slower than the default below on the 1024-case
        # sweep (profiles/r2i_sweep_*: 2.35e9 against 3.15e9 dof-updates/s): the cases do not fit the caches together, so
        # everything outside the solves streams from HBM at one CTA per SM.  Opt-in.
        pc = _lib.loop_patch_cells(len(mesh[1]))
        if pc:
            patch_cells, solver_flags = pc, _lib.FLAGS_LOOP
    elif auto and len(mesh[1]) > 512:
        # default: all cases side by side in the bulk-copy pipelined mass solves (flags bits 8, 9; patches of <= 1024 cells)
        patch_cells, solver_flags = 1024, 256 | 512
    if patch_cells is None:
        patch_cells = max(64, min(2048, len(mesh[1])))
    m = Mesh(V, C, patch_cells=patch_cells, cell_group=group, n_groups=B)
    m.set_boundary(tag_boundary_edges(m, domainShape, xLength, yLength))
    prm = _lib.PhwParams()
    prm.scheme, prm.interconnection = _lib.SCHEMES[timeIntScheme], 1
    prm.dt, prm.K_wave, prm.rho = tFinal / numSteps, 1.0, 1.0
    for k, v in DEFAULT_LUMPED.items():
        setattr(prm, k, v)
    if B == 1:   # a single case (e.g. a one-case shard of a sweep split over GPUs) has no groups: its physics goes through phw_params
        prm.K_wave, prm.rho = float(cases[0].get('K_wave', 1.0)), float(cases[0].get('rho', 1.0))
        for k in LUMPED_KEYS:
            setattr(prm, k, float(cases[0].get(k, DEFAULT_LUMPED[k])))
    prm.input_kind, prm.in_amp, prm.in_omega, prm.in_t_stop = _lib.IN_SINE_WINDOW, 5.0, 8 * np.pi, 0.25     # wave_2D.py:433
    prm.tol, prm.max_iter, prm.extrap_order, prm.flags = float(tol), 400, int(extrap_order), int(solver_flags)
    try:
        solver = Solver(m, prm, device=device)
    except _lib.PhwError as ex:
        if not (auto and 'shared memory' in str(ex)):
            raise
        # the two stages of the pipelined solves do not fit for this mesh: lean kernels on 2048-cell patches
        m.close()
        m = Mesh(V, C, patch_cells=max(64, min(2048, len(mesh[1]))), cell_group=group, n_groups=B)
        m.set_boundary(tag_boundary_edges(m, domainShape, xLength, yLength))
        prm.flags = 0
        solver = Solver(m, prm, device=device)
    K = np.array([c.get('K_wave', 1.0) for c in cases], dtype=np.float64)
    rho = np.array([c.get('rho', 1.0) for c in cases], dtype=np.float64)
    lum = np.array([[c.get(k, DEFAULT_LUMPED[k]) for k in LUMPED_KEYS] for c in cases], dtype=np.float64)
    if B > 1:
        solver.set_group_params(K, rho, lum)
    H = solver.run(numSteps)
    if B == 1:
        H = H[None]
    numCells = len(mesh[1])
    if return_solver:
        return H, numCells, solver
    solver.close()
    m.close()
    return H, numCells


def _sweep_case_fastest(tFinal, numSteps, mesh, xLength, yLength, cases, domainShape, timeIntScheme, device, tol, extrap_order, return_solver):
    """The sweep on ONE set of assembled P1 x RT1 operators (host assembly of fem_ho.py at order (1, 1): M_p, M_q, D = C - N, b_L and
    the element-wise bounds of the Jacobi-scaled mass spectra), all cases side by side in [n_dof][n_cases] vectors."""
    from .fem_ho import HighOrderSpaces
    from .solver import BatchedSweepSolver, tag_space_boundary
    S = HighOrderSpaces(mesh[0], mesh[1], 1, 1)
    ops = S.assemble(tag_space_boundary(S, domainShape, xLength, yLength))                 # facet markers of wave_2D.py:226-280
    bp, bq = S.jacobi_bounds(ops)
    B = len(cases)
    prm = _lib.PhwParams()
    prm.scheme, prm.interconnection = _lib.SCHEMES[timeIntScheme], 1
    prm.dt, prm.K_wave, prm.rho = tFinal / numSteps, 1.0, 1.0
    for k, v in DEFAULT_LUMPED.items():
        setattr(prm, k, v)
    prm.input_kind, prm.in_amp, prm.in_omega, prm.in_t_stop = _lib.IN_SINE_WINDOW, 5.0, 8 * np.pi, 0.25     # wave_2D.py:433
    prm.tol, prm.max_iter, prm.extrap_order, prm.flags = float(tol), 400, int(extrap_order), 0
    solver = BatchedSweepSolver(S, ops, prm, bp, bq, B, device=device)
    K = np.array([c.get('K_wave', 1.0) for c in cases], dtype=np.float64)
    rho = np.array([c.get('rho', 1.0) for c in cases], dtype=np.float64)
    lum = np.array([[c.get(k, DEFAULT_LUMPED[k]) for k in LUMPED_KEYS] for c in cases], dtype=np.float64)
    solver.set_group_params(K, rho, lum)
    H = solver.run(numSteps)
    numCells = len(mesh[1])
    if return_solver:
        return H, numCells, solver
    solver.close()
    return H, numCells
