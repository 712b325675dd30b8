"""GPU parity of the single-launch whole-loop kernel (csrc/kernels_loop.cuh; phw_params.flags bit 15): the complete time loop of
wave_2D.py:787-982 - right-hand sides, Chebyshev solves, port flux, lumped ODE / controller, energies, H_array rows - runs inside
ONE kernel, the CTAs of a case forming one thread-block cluster (or one co-resident grid).  It is the default of wave_2D_solve and
wave_sweep_solve for cache-resident meshes, i.e. for every case the reference itself runs; every scheme and variant is compared
with the CPU oracle exactly like the multi-launch path (tests/test_gpu_parity.py), and with the multi-launch path itself."""
import numpy as np
import pytest

from oracle.wave_2D_oracle import wave_2D_solve_oracle
from porthamiltonian_fem_b200 import _lib, wave_2D_solve
from porthamiltonian_fem_b200.mesh import rectangle_structured, rectangle_delaunay, square_one_channel
from porthamiltonian_fem_b200.solver import Mesh, Solver, tag_boundary_edges

pytestmark = pytest.mark.gpu
K_WAVE, RHO = 3.0, 2.0
TOL_STATE = 1e-10
PATH_LOOP = 256


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


CASES = [
    ('EE_weak', 34, 8, 120, 0.06, dict(timeIntScheme='EE')),
    ('SE_weak', 34, 8, 120, 0.06, dict(timeIntScheme='SE')),
    ('EH_weak', 34, 8, 120, 0.06, dict(timeIntScheme='EH')),
    ('SV_weak', 34, 8, 120, 0.06, dict(timeIntScheme='SV')),
    ('IE_weak', 34, 8, 120, 0.06, dict(timeIntScheme='IE')),
    ('SM_weak', 34, 8, 120, 0.06, dict(timeIntScheme='SM')),
    ('SV_strong', 34, 8, 120, 0.06, dict(timeIntScheme='SV', dirichletImp='strong')),
    ('SE_strong', 34, 8, 120, 0.06, dict(timeIntScheme='SE', dirichletImp='strong')),
    ('EH_strong', 34, 8, 120, 0.06, dict(timeIntScheme='EH', dirichletImp='strong')),
    ('SM_strong', 34, 8, 120, 0.06, dict(timeIntScheme='SM', dirichletImp='strong')),
    ('SE_IC', 52, 13, 100, 0.1, dict(timeIntScheme='SE', interConnection='IC')),
    ('SV_IC', 52, 13, 100, 0.1, dict(timeIntScheme='SV', interConnection='IC')),
    ('SM_IC', 52, 13, 100, 0.1, dict(timeIntScheme='SM', interConnection='IC')),
    ('SM_casimir', 40, 10, 100, 0.1, dict(timeIntScheme='SM', controlType='casimir')),
    ('SM_passivity', 40, 10, 150, 0.15, dict(timeIntScheme='SM', interConnection='IC', controlType='passivity',
                                             input_stop_t=0.02, control_start_t=0.06)),
    ('SE_passivity', 40, 10, 150, 0.15, dict(timeIntScheme='SE', interConnection='IC', controlType='passivity',
                                             input_stop_t=0.02, control_start_t=0.06)),
]


def check(Hg, sg, Ho, so, ic):
    assert sg['stats']['kernel_paths'] & PATH_LOOP, 'the whole-loop kernel did not run: kernel_paths = %d' % sg['stats']['kernel_paths']
    assert sg['stats']['kernel_launches'] == 1
    assert Hg.shape == Ho.shape and np.array_equal(Hg[:, 0], Ho[:, 0])
    assert rel(sg['p'], so['p']) < TOL_STATE, rel(sg['p'], so['p'])
    assert rel(sg['q'], so['q']) < TOL_STATE, rel(sg['q'], so['q'])
    if ic:
        assert rel(sg['x'], so['x']) < TOL_STATE
    scale = np.abs(Ho[:, [1, 4, 5, 6, 7]]).max() + 1e-300
    for col in range(1, Ho.shape[1]):
        assert np.abs(Hg[:, col] - Ho[:, col]).max() < 1e-10 * max(scale, np.abs(Ho[:, col]).max()), col
    assert sg['stats']['max_rel_residual'] <= 1e-13


@pytest.mark.parametrize('name,Nx,Ny,steps,tF,kw', CASES)
def test_whole_loop_kernel_parity(name, Nx, Ny, steps, tF, kw):
    """drop-in default on a small mesh (one cluster of 1 - 2 CTAs here)"""
    mesh = rectangle_structured(Nx, Ny, 1.0, 0.25)
    Ho, nco, so = wave_2D_solve_oracle(tF, steps, mesh, 1.0, 0.25, K_wave=K_WAVE, rho=RHO, return_state=True, **kw)
    Hg, ncg, sg = wave_2D_solve(tF, steps, '/tmp', 0, 1.0, 0.25, K_wave=K_WAVE, rho=RHO, mesh=mesh, return_state=True, verbose=False, **kw)
    assert ncg == nco
    check(Hg, sg, Ho, so, kw.get('interConnection') == 'IC')


# cluster sizes 1 .. 16 and the co-resident-grid mode (a patch count that is no cluster size); patch_cells chosen explicitly
LAYOUTS = [
    ('cluster_1', lambda: rectangle_structured(20, 6), 'R', 1.0, 0.25, 4096),
    ('cluster_2', lambda: rectangle_structured(40, 10), 'R', 1.0, 0.25, 400),
    ('cluster_4', lambda: rectangle_delaunay(60, 15, seed=5), 'R', 1.0, 0.25, 450),
    ('cluster_8', lambda: rectangle_structured(136, 34), 'R', 1.0, 0.25, 1156),
    ('cluster_16', lambda: rectangle_structured(136, 34), 'R', 1.0, 0.25, 578),
    ('cluster_8_s1c', lambda: square_one_channel(60), 'S_1C', 1.0, 0.1, None),
    ('grid_24', lambda: rectangle_structured(96, 24), 'R', 1.0, 0.25, 192),
    ('grid_36', lambda: rectangle_delaunay(100, 25, seed=2), 'R', 1.0, 0.25, 140),
]


@pytest.mark.parametrize('name,make,shape,xL,yL,pc', LAYOUTS)
@pytest.mark.parametrize('kw', [dict(timeIntScheme='SV', interConnection='IC'), dict(timeIntScheme='SM', interConnection='IC'),
                                dict(timeIntScheme='SE')], ids=['SV_IC', 'SM_IC', 'SE'])
def test_whole_loop_kernel_cluster_sizes(name, make, shape, xL, yL, pc, kw):
    mesh = make()
    steps, tF = 40, 0.02
    if pc is None:
        pc = _lib.loop_patch_cells(len(mesh[1]))
        assert pc is not None
    Ho, _, so = wave_2D_solve_oracle(tF, steps, mesh, xL, yL, domainShape=shape, K_wave=K_WAVE, rho=RHO, return_state=True, **kw)
    Hg, _, sg = wave_2D_solve(tF, steps, '/tmp', 0, xL, yL, domainShape=shape, K_wave=K_WAVE, rho=RHO, mesh=mesh, patch_cells=pc,
                              solver_flags=_lib.FLAGS_LOOP, return_state=True, verbose=False, **kw)
    assert sg['mesh'].n_patches == int(name.split('_')[1])
    check(Hg, sg, Ho, so, 'interConnection' in kw)


def test_whole_loop_kernel_restart_and_mixed_paths():
    """two runs of 30 steps continue exactly like one run of 60 (history ring, first-same-as-last flag and iterate buffers are
    carried across launches), and a run that alternates between the whole-loop kernel and kernel-level entry points stays exact"""
    v, c = rectangle_delaunay(48, 12, seed=3)
    out = []
    for chunks in ((60,), (30, 30), (1, 2, 57)):
        m = Mesh(v, c, patch_cells=_lib.loop_patch_cells(len(c)))
        m.set_boundary(tag_boundary_edges(m, 'R', 1.0, 0.25))
        prm = _lib.PhwParams()
        prm.scheme, prm.interconnection, prm.dt, prm.K_wave, prm.rho = _lib.SCHEMES['SV'], 1, 5e-4, K_WAVE, RHO
        prm.Bl_em, prm.K_em, prm.m_em, prm.L_em = 5.0, 5.0, 0.15, 0.1
        prm.input_kind, prm.in_amp, prm.in_omega, prm.in_t_stop = _lib.IN_SINE_WINDOW, 5.0, 8 * np.pi, 0.25
        prm.tol, prm.max_iter, prm.extrap_order, prm.flags = 1e-13, 400, 3, _lib.FLAGS_LOOP
        s = Solver(m, prm)
        H = np.concatenate([s.run(n) for n in chunks])
        assert s.stats()['kernel_paths'] & PATH_LOOP
        out.append((H, s.get_state()))
    for H, st in out[1:]:
        assert np.array_equal(H[:, 0], out[0][0][:, 0])
        assert np.abs(H[:, 1:] - out[0][0][:, 1:]).max() < 1e-11 * np.abs(out[0][0][:, 1:]).max()
        for a, b in zip(st, out[0][1]):
            assert rel(a, b) < 1e-11


def test_whole_loop_kernel_matches_the_multi_launch_path():
    mesh = rectangle_structured(52, 13)
    for kw in (dict(timeIntScheme='SV', interConnection='IC'), dict(timeIntScheme='SM', controlType='casimir'), dict(timeIntScheme='EH')):
        a = wave_2D_solve(0.05, 50, '/tmp', 0, 1.0, 0.25, K_wave=K_WAVE, rho=RHO, mesh=mesh, return_state=True, verbose=False, **kw)
        b = wave_2D_solve(0.05, 50, '/tmp', 0, 1.0, 0.25, K_wave=K_WAVE, rho=RHO, mesh=mesh, return_state=True, verbose=False,
                          solver_flags=0, patch_cells=128, **kw)
        assert a[2]['stats']['kernel_paths'] & PATH_LOOP and not (b[2]['stats']['kernel_paths'] & PATH_LOOP)
        assert b[2]['stats']['kernel_launches'] > 50
        assert rel(a[2]['p'], b[2]['p']) < 1e-11 and rel(a[2]['q'], b[2]['q']) < 1e-11
        assert np.abs(a[0][:, 1:] - b[0][:, 1:]).max() < 1e-11 * np.abs(b[0][:, 1:]).max()


@pytest.mark.parametrize('loop', [True, False])
def test_batched_sweep_runs_one_cluster_per_case(loop, monkeypatch):
    """config 5 in small.  loop (PHW_SWEEP_LOOP=1): every case is advanced by its own cluster inside one launch; default: all
    cases side by side in the pipelined mass solves.  Each H_array equals what wave_2D_solve's oracle gives for that case alone;
    a one-case sweep takes its physics from the case"""
    from porthamiltonian_fem_b200.sweep import wave_sweep_solve, sweep_cases, LUMPED_KEYS, DEFAULT_LUMPED
    if loop:
        monkeypatch.setenv('PHW_SWEEP_LOOP', '1')
    else:
        monkeypatch.delenv('PHW_SWEEP_LOOP', raising=False)
    mesh = rectangle_structured(52, 13, 1.0, 0.25)             # 1352 cells -> 2 patches per case
    cases = sweep_cases(7)
    H, nc, solver = wave_sweep_solve(0.04, 40, 0, 1.0, 0.25, cases, mesh=mesh, return_solver=True)
    st = solver.stats()
    if loop:
        assert st['kernel_paths'] & PATH_LOOP and st['kernel_launches'] == 1
    else:
        assert st['kernel_paths'] & 12 == 12 and not st['kernel_paths'] & PATH_LOOP      # PHW_PATH_PIPE_P | PHW_PATH_PIPE_Q
    for i, c in enumerate(cases):
        lum = {k: c.get(k, DEFAULT_LUMPED[k]) for k in LUMPED_KEYS}
        Ho, _, _ = wave_2D_solve_oracle(0.04, 40, mesh, 1.0, 0.25, timeIntScheme='SV', interConnection='IC',
                                        K_wave=c['K_wave'], rho=c['rho'], lumped=lum, return_state=True)
        assert np.abs(H[i][:, 1:] - Ho[:, 1:]).max() < 1e-10 * np.abs(Ho[:, 1:]).max(), i
    H1, _ = wave_sweep_solve(0.04, 40, 0, 1.0, 0.25, cases[3:4], mesh=mesh)
    assert np.abs(H1[0][:, 1:] - H[3][:, 1:]).max() < 1e-11 * np.abs(H[3][:, 1:]).max()


@pytest.mark.parametrize('scheme,n_cases', [('SV', 40), ('SE', 33)])
def test_sweep_with_the_case_index_fastest_matches_the_oracle_case_by_case(scheme, n_cases, monkeypatch):
    """config 5, layout of SURVEY K9 (default for >= 32 cases): ONE copy of the assembled operators, state [n_dof][n_cases]
    (csrc/kernels_sweep.cuh).  Every case's H_array equals what the oracle gives for that case alone; the final fields and lumped
    states of two cases too; the replicated-mesh layout (PHW_SWEEP_GROUPED=1) gives the same traces."""
    from porthamiltonian_fem_b200.sweep import wave_sweep_solve, sweep_cases, LUMPED_KEYS, DEFAULT_LUMPED
    monkeypatch.delenv('PHW_SWEEP_GROUPED', raising=False)
    monkeypatch.delenv('PHW_SWEEP_LOOP', raising=False)
    mesh = rectangle_structured(20, 5, 1.0, 0.25)
    cases = sweep_cases(n_cases, seed=7)
    steps, tF = 40, 0.04
    H, nc, solver = wave_sweep_solve(tF, steps, 0, 1.0, 0.25, cases, mesh=mesh, timeIntScheme=scheme, return_solver=True)
    st = solver.stats()
    assert st['kernel_paths'] & 4096 and H.shape == (n_cases, steps, 8) and nc == len(mesh[1])       # PHW_PATH_SWEEP_CSR
    assert st['max_rel_residual'] <= 1e-13
    P, Q, X = solver.get_case_states()
    solver.close()
    for i in list(range(0, n_cases, 5)) + [n_cases - 1]:
        c = cases[i]
        lum = {k: c.get(k, DEFAULT_LUMPED[k]) for k in LUMPED_KEYS}
        Ho, _, so = wave_2D_solve_oracle(tF, steps, mesh, 1.0, 0.25, timeIntScheme=scheme, interConnection='IC',
                                         K_wave=c['K_wave'], rho=c['rho'], lumped=lum, return_state=True)
        assert np.array_equal(H[i][:, 0], Ho[:, 0])
        assert np.abs(H[i][:, 1:] - Ho[:, 1:]).max() < 1e-10 * np.abs(Ho[:, 1:]).max(), i
        if i in (0, n_cases - 1):
            assert np.linalg.norm(P[i] - so['p']) < 1e-10 * np.linalg.norm(so['p'])
            assert np.linalg.norm(Q[i] - so['q']) < 1e-10 * np.linalg.norm(so['q'])
            assert np.abs(X[i] - so['x']).max() < 1e-10 * max(1.0, np.abs(so['x']).max())
    monkeypatch.setenv('PHW_SWEEP_GROUPED', '1')
    Hg, _ = wave_sweep_solve(tF, steps, 0, 1.0, 0.25, cases, mesh=mesh, timeIntScheme=scheme)
    assert np.abs(Hg[:, :, 1:] - H[:, :, 1:]).max() < 1e-10 * np.abs(H[:, :, 1:]).max()
