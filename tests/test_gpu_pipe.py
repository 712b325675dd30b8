"""GPU parity tests of the bulk-copy pipelined mass solves (csrc/kernels_pipe.cuh; phw_params.flags bits 8 and 9):
same Chebyshev iteration as the lean kernels, operands staged by cp.async.bulk one patch ahead.  Compared with the
CPU oracle's sparse direct solves and with the lean kernels, through the C ABI.  Meshes are chosen so that every CTA
walks several patches (the steady state of the two-stage pipeline) and so that the owned dof ranges start at every
residue mod 2 / mod 4 (the unaligned heads and tails of the bulk copies)."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from oracle import fem
from oracle.wave_2D_oracle import wave_2D_solve_oracle
from porthamiltonian_fem_b200 import _lib, wave_2D_solve
from porthamiltonian_fem_b200.mesh import rectangle_structured, rectangle_delaunay, square_one_channel
from test_gpu_parity import make_solver, rel, K_WAVE, RHO, TOL_STATE

pytestmark = pytest.mark.gpu
PIPE = 256 | 512
ENERGY = 2048
RHS = 4096
PATH_PIPE_P, PATH_PIPE_Q, PATH_PIPE_ENERGY, PATH_PIPE_RHS = 4, 8, 16, 32
ALL_PIPE = PIPE | ENERGY | RHS

MESHES = [
    ('structured_34x8_pc64', lambda: rectangle_structured(34, 8), 'R', 1.0, 0.25, 64),
    ('structured_400x100_pc64_many_patches', lambda: rectangle_structured(400, 100), 'R', 1.0, 0.25, 64),
    ('structured_300x75_pc1000', lambda: rectangle_structured(300, 75), 'R', 1.0, 0.25, 1000),
    ('delaunay_240x60_pc100', lambda: rectangle_delaunay(240, 60, seed=3), 'R', 1.0, 0.25, 100),
    ('s1c_40_pc96', lambda: square_one_channel(40), 'S_1C', 1.0, 0.1, 96),
    ('single_patch', lambda: rectangle_structured(20, 10), 'R', 1.0, 0.25, 8192),
    ('tiny_2x1', lambda: rectangle_structured(2, 1), 'R', 1.0, 0.25, 16),
]


@pytest.mark.parametrize('name,make,shape,xL,yL,pc', MESHES)
def test_pipelined_mass_solves_match_oracle_and_lean_kernels(name, make, shape, xL, yL, pc):
    v, c = make()
    m, s = make_solver(v, c, shape, xL, yL, pc, flags=PIPE)
    topo = fem.Topology(v, c)
    ops = fem.assemble_wave_operators(topo, fem.tag_boundary(topo, shape, xL, yL))
    rng = np.random.default_rng(17)
    bp = rng.standard_normal(topo.n_vertices)
    bq = rng.standard_normal(topo.n_edges)
    xp, itp = s.mass_solve(_lib.PHW_P, bp)
    xq, itq = s.mass_solve(_lib.PHW_Q, bq)
    assert s.stats()['kernel_paths'] & (PATH_PIPE_P | PATH_PIPE_Q) == (PATH_PIPE_P | PATH_PIPE_Q)
    assert rel(xp, spla.spsolve(ops['M_p'].tocsc(), bp)) < 1e-11
    assert rel(xq, spla.spsolve(ops['M_q'].tocsc(), bq)) < 1e-11
    m1, s1 = make_solver(v, c, shape, xL, yL, pc)
    xp1, itp1 = s1.mass_solve(_lib.PHW_P, bp)
    xq1, itq1 = s1.mass_solve(_lib.PHW_Q, bq)
    assert abs(itp - itp1) <= 1 and abs(itq - itq1) <= 1      # same iteration; the stop test sums the norms in another order
    assert rel(xp, xp1) < 1e-12 and rel(xq, xq1) < 1e-12
    # repeated solves reuse the stages and mbarriers with alternating parities
    for _ in range(3):
        xq2, _ = s.mass_solve(_lib.PHW_Q, bq)
        assert rel(xq2, xq) < 1e-12
    z, _ = s.mass_solve(_lib.PHW_Q, np.zeros(topo.n_edges))
    assert np.all(z == 0.0)


@pytest.mark.parametrize('name,make,shape,xL,yL,pc', MESHES)
def test_pipelined_energy_reduction_matches_oracle(name, make, shape, xL, yL, pc):
    """flags bit 11: both quadratic forms of H_wave in one pipelined pass over the own cells of every patch"""
    v, c = make()
    m, s = make_solver(v, c, shape, xL, yL, pc, flags=ENERGY)
    topo = fem.Topology(v, c)
    ops = fem.assemble_wave_operators(topo, fem.tag_boundary(topo, shape, xL, yL))
    rng = np.random.default_rng(23)
    for _ in range(2):
        p = rng.standard_normal(topo.n_vertices)
        q = rng.standard_normal(topo.n_edges)
        s.set_state(p=p, q=q, x3=np.zeros(3))
        Hp, Hq = s.energy_parts()
        assert abs(Hp - 0.5 * p @ (ops['M_p'] @ p) / RHO) < 1e-12 * abs(Hp)
        assert abs(Hq - 0.5 * K_WAVE * q @ (ops['M_q'] @ q)) < 1e-12 * abs(Hq)
    assert s.stats()['kernel_paths'] & PATH_PIPE_ENERGY


@pytest.mark.parametrize('name,make,shape,xL,yL,pc', MESHES)
def test_pipelined_right_hand_sides_match_oracle(name, make, shape, xL, yL, pc):
    """flags bit 12: D p and D^T q (assemble(b), wave_2D.py:814,849) through the pipelined row kernels"""
    v, c = make()
    m, s = make_solver(v, c, shape, xL, yL, pc, flags=RHS)
    topo = fem.Topology(v, c)
    ops = fem.assemble_wave_operators(topo, fem.tag_boundary(topo, shape, xL, yL))
    rng = np.random.default_rng(29)
    for _ in range(2):
        p = rng.standard_normal(topo.n_vertices)
        q = rng.standard_normal(topo.n_edges)
        assert rel(s.apply_D(p), ops['D'] @ p) < 1e-13
        assert rel(s.apply_DT(q), ops['D'].T @ q) < 1e-13
    assert s.stats()['kernel_paths'] & PATH_PIPE_RHS
    m0, s0 = make_solver(v, c, shape, xL, yL, pc)
    assert rel(s.apply_D(p), s0.apply_D(p)) < 1e-14 and rel(s.apply_DT(q), s0.apply_DT(q)) < 1e-14   # same summation order


def test_pipelined_right_hand_sides_strong_form():
    mesh = rectangle_delaunay(120, 30, seed=6)
    kw = dict(timeIntScheme='SV', dirichletImp='strong')
    Ho, _, so = wave_2D_solve_oracle(0.03, 60, mesh, 1.0, 0.25, K_wave=K_WAVE, rho=RHO, return_state=True, **kw)
    Hg, _, sg = wave_2D_solve(0.03, 60, '/tmp', 0, 1.0, 0.25, K_wave=K_WAVE, rho=RHO, mesh=mesh, patch_cells=48,
                              return_state=True, verbose=False, solver_flags=RHS | ENERGY | 256, **kw)
    assert sg['stats']['kernel_paths'] & PATH_PIPE_RHS
    assert rel(sg['p'], so['p']) < TOL_STATE and rel(sg['q'], so['q']) < TOL_STATE
    assert np.abs(Hg[:, 1:] - Ho[:, 1:]).max() < 1e-10 * np.abs(Ho[:, 1:]).max()


@pytest.mark.parametrize('mq_diag', ['1', '0'])
def test_pipelined_mq_solve_with_and_without_streamed_jacobi_weights(mq_diag, monkeypatch):
    """PHW_MQ_PIPE_DIAG (A/B knob, read once per process - so this test runs the solver in a child process)"""
    import subprocess, sys, os, textwrap
    code = textwrap.dedent("""
        import numpy as np, scipy.sparse.linalg as spla
        from oracle import fem
        from porthamiltonian_fem_b200 import _lib
        from porthamiltonian_fem_b200.mesh import rectangle_delaunay
        from test_gpu_parity import make_solver, rel
        v, c = rectangle_delaunay(200, 50, seed=9)
        m, s = make_solver(v, c, 'R', 1.0, 0.25, 80, flags=256)
        topo = fem.Topology(v, c)
        ops = fem.assemble_wave_operators(topo, fem.tag_boundary(topo, 'R', 1.0, 0.25))
        bq = np.random.default_rng(5).standard_normal(topo.n_edges)
        xq, it = s.mass_solve(_lib.PHW_Q, bq)
        assert s.stats()['kernel_paths'] & 8
        assert rel(xq, spla.spsolve(ops['M_q'].tocsc(), bq)) < 1e-11
        print('ok', it)
    """)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PHW_MQ_PIPE_DIAG=mq_diag, PYTHONPATH=root + os.pathsep + os.path.join(root, 'tests'))
    r = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and 'ok' in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize('scheme,ic', [('SE', None), ('SV', 'IC'), ('EH', None)])
def test_pipelined_time_loop_parity(scheme, ic):
    mesh = rectangle_delaunay(120, 30, seed=5)
    steps, tF = 60, 0.03
    kw = dict(timeIntScheme=scheme, interConnection=ic)
    Ho, _, so = wave_2D_solve_oracle(tF, steps, mesh, 1.0, 0.25, K_wave=K_WAVE, rho=RHO, return_state=True, **kw)
    Hg, _, sg = wave_2D_solve(tF, steps, '/tmp', 0, 1.0, 0.25, K_wave=K_WAVE, rho=RHO, mesh=mesh, patch_cells=48,
                              return_state=True, verbose=False, solver_flags=ALL_PIPE, **kw)
    want = PATH_PIPE_P | PATH_PIPE_Q | PATH_PIPE_ENERGY | PATH_PIPE_RHS
    assert sg['stats']['kernel_paths'] & want == want
    assert rel(sg['p'], so['p']) < TOL_STATE and rel(sg['q'], so['q']) < TOL_STATE
    assert np.abs(Hg[:, 1:] - Ho[:, 1:]).max() < 1e-10 * np.abs(Ho[:, 1:]).max()
    assert sg['stats']['max_rel_residual'] <= 1e-13


def test_pipelined_solve_rejects_patches_that_do_not_fit():
    v, c = rectangle_structured(200, 50)
    with pytest.raises(Exception):
        make_solver(v, c, 'R', 1.0, 0.25, 4096, flags=256)


# ---- flags bit 10: the generic cooperative kernels synchronise the grid with the one-atomic-per-CTA barrier ----
GBAR = 1024
GBAR_CASES = [
    ('SM_IC', 52, 13, 100, 0.1, dict(timeIntScheme='SM', interConnection='IC')),
    ('IE_wave', 34, 8, 120, 0.06, dict(timeIntScheme='IE')),
    ('SE_strong', 34, 8, 120, 0.06, dict(timeIntScheme='SE', dirichletImp='strong')),
    ('SM_strong', 34, 8, 120, 0.06, dict(timeIntScheme='SM', dirichletImp='strong')),
    ('SM_casimir', 40, 10, 100, 0.1, dict(timeIntScheme='SM', controlType='casimir')),
    ('SE_P2RT2', 20, 5, 100, 0.03, dict(timeIntScheme='SE', basis_order=(2, 2))),
]


@pytest.mark.parametrize('name,Nx,Ny,steps,tF,kw', GBAR_CASES)
def test_custom_grid_barrier_in_generic_kernels(name, Nx, Ny, steps, tF, kw):
    mesh = rectangle_structured(Nx, Ny, 1.0, 0.25)
    kw = dict(kw)
    flags = kw.pop('solver_flags', GBAR)
    Ho, _, so = wave_2D_solve_oracle(tF, steps, mesh, 1.0, 0.25, K_wave=K_WAVE, rho=RHO, return_state=True, **kw)
    extra = {} if 'basis_order' in kw else dict(patch_cells=64)
    Hg, _, sg = wave_2D_solve(tF, steps, '/tmp', 0, 1.0, 0.25, K_wave=K_WAVE, rho=RHO, mesh=mesh, return_state=True,
                              verbose=False, solver_flags=flags, **extra, **kw)
    assert np.array_equal(Hg[:, 0], Ho[:, 0])
    assert rel(sg['p'], so['p']) < TOL_STATE and rel(sg['q'], so['q']) < TOL_STATE
    assert np.abs(Hg[:, 1:] - Ho[:, 1:]).max() < 1e-10 * np.abs(Ho[:, 1:]).max()


def test_custom_grid_barrier_batched_sweep():
    from porthamiltonian_fem_b200.sweep import wave_sweep_solve, sweep_cases, LUMPED_KEYS, DEFAULT_LUMPED
    mesh = rectangle_structured(26, 6, 1.0, 0.25)
    cases = sweep_cases(5)
    H, nc = wave_sweep_solve(0.04, 40, 0, 1.0, 0.25, cases, mesh=mesh, patch_cells=96, solver_flags=GBAR | 128)
    for i, c in enumerate(cases):
        lum = {k: c.get(k, DEFAULT_LUMPED[k]) for k in LUMPED_KEYS}
        Ho, _, _ = wave_2D_solve_oracle(0.04, 40, mesh, 1.0, 0.25, timeIntScheme='SV', interConnection='IC',
                                        K_wave=c['K_wave'], rho=c['rho'], lumped=lum, return_state=True)
        assert np.abs(H[i][:, 1:] - Ho[:, 1:]).max() < 1e-10 * np.abs(Ho[:, 1:]).max(), i


def test_drop_in_default_on_a_large_mesh_runs_the_pipelined_kernels():
    """wave_2D_solve without solver_flags on >= 200 000 cells: 1024-cell patches + all pipelined kernels; parity as everywhere"""
    mesh = rectangle_structured(640, 160)                      # 204 800 cells, 410 k dofs
    steps, tF = 8, 0.0016
    Ho, _, so = wave_2D_solve_oracle(tF, steps, mesh, 1.0, 0.25, K_wave=K_WAVE, rho=RHO, return_state=True, timeIntScheme='SE')
    Hg, nc, sg = wave_2D_solve(tF, steps, '/tmp', 0, 1.0, 0.25, K_wave=K_WAVE, rho=RHO, mesh=mesh, timeIntScheme='SE',
                               return_state=True, verbose=False)
    want = PATH_PIPE_P | PATH_PIPE_Q | PATH_PIPE_ENERGY | PATH_PIPE_RHS
    assert nc == 204800 and sg['stats']['kernel_paths'] & want == want
    assert rel(sg['p'], so['p']) < TOL_STATE and rel(sg['q'], so['q']) < TOL_STATE
    assert np.abs(Hg[:, 1:] - Ho[:, 1:]).max() < 1e-10 * np.abs(Ho[:, 1:]).max()
