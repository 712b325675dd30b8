"""GPU parity tests (run on the B200 box): every call goes through the C ABI of libphwave.so and is compared
with the CPU oracle on identical meshes and seeded inputs.

Tolerances (BASELINE.json north_star): integer indexing bit-exact (tests/test_layout_cpu.py), floating point
state (p, q, lumped x) relative L2 error <= 1e-10 after N steps, H_array traces to the same level."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from oracle import fem
from oracle.wave_2D_oracle import wave_2D_solve_oracle
from porthamiltonian_fem_b200 import _lib, wave_2D_solve
from porthamiltonian_fem_b200.mesh import rectangle_structured, rectangle_delaunay, square_one_channel
from porthamiltonian_fem_b200.solver import Mesh, Solver, tag_boundary_edges

pytestmark = pytest.mark.gpu
K_WAVE, RHO = 3.0, 2.0
TOL_STATE = 1e-10


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def make_solver(v, c, shape='R', xL=1.0, yL=0.25, pc=128, scheme='SE', **extra):
    m = Mesh(v, c, patch_cells=pc)
    m.set_boundary(tag_boundary_edges(m, shape, xL, yL))
    prm = _lib.PhwParams()
    prm.scheme, prm.dt, prm.K_wave, prm.rho = _lib.SCHEMES[scheme], 1e-3, K_WAVE, RHO
    prm.input_kind, prm.in_amp, prm.in_omega, prm.in_t_stop = _lib.IN_SINE_WINDOW, 10.0, 8 * np.pi, 0.25
    prm.tol, prm.max_iter = 1e-13, 200
    for k, val in extra.items():
        setattr(prm, k, val)
    return m, Solver(m, prm)


MESHES = [
    ('structured_34x8_pc64', lambda: rectangle_structured(34, 8), 'R', 1.0, 0.25, 64),
    ('structured_136x34_pc2048', lambda: rectangle_structured(136, 34), 'R', 1.0, 0.25, 2048),
    ('structured_60x40_single_patch', lambda: rectangle_structured(60, 40), 'R', 1.0, 0.25, 8192),
    ('delaunay_80x20_pc300', lambda: rectangle_delaunay(80, 20, seed=1), 'R', 1.0, 0.25, 300),
    ('s1c_20_pc256', lambda: square_one_channel(20), 'S_1C', 1.0, 0.1, 256),
    ('tiny_2x1', lambda: rectangle_structured(2, 1), 'R', 1.0, 0.25, 16),
]


@pytest.mark.parametrize('name,make,shape,xL,yL,pc', MESHES)
def test_operator_parity(name, make, shape, xL, yL, pc):
    v, c = make()
    m, s = make_solver(v, c, shape, xL, yL, pc)
    topo = fem.Topology(v, c)
    ops = fem.assemble_wave_operators(topo, fem.tag_boundary(topo, shape, xL, yL))
    rng = np.random.default_rng(7)
    p = rng.standard_normal(topo.n_vertices)
    q = rng.standard_normal(topo.n_edges)
    assert rel(s.apply_D(p), ops['D'] @ p) < 1e-13
    assert rel(s.apply_DT(q), ops['D'].T @ q) < 1e-13
    assert rel(s.mass_matvec(_lib.PHW_P, p), ops['M_p'] @ p) < 1e-13
    assert rel(s.mass_matvec(_lib.PHW_Q, q), ops['M_q'] @ q) < 1e-13
    # skew-adjointness of the interconnection operator: q.(D p) == p.(D^T q)
    assert abs(q @ s.apply_D(p) - p @ s.apply_DT(q)) < 1e-11 * np.linalg.norm(p) * np.linalg.norm(q)
    # energy and port flux reductions
    s.set_state(p=p, q=q, x3=np.zeros(3))
    Hp, Hq = s.energy_parts()
    assert abs(Hp - 0.5 * p @ (ops['M_p'] @ p) / RHO) < 1e-12 * abs(Hp)
    assert abs(Hq - 0.5 * K_WAVE * q @ (ops['M_q'] @ q)) < 1e-12 * abs(Hq)
    assert abs(s.port_flux() - ops['b_L'] @ q) < 1e-12 * max(1.0, abs(ops['b_L'] @ q))
    pp, qq, _ = s.get_state()
    assert np.array_equal(pp, p) and np.array_equal(qq, q)      # permutation round trip is exact


@pytest.mark.parametrize('name,make,shape,xL,yL,pc', MESHES)
def test_mass_solve_parity(name, make, shape, xL, yL, pc):
    v, c = make()
    m, s = make_solver(v, c, shape, xL, yL, pc)
    topo = fem.Topology(v, c)
    ops = fem.assemble_wave_operators(topo, fem.tag_boundary(topo, shape, xL, yL))
    rng = np.random.default_rng(11)
    bp = rng.standard_normal(topo.n_vertices)
    bq = rng.standard_normal(topo.n_edges)
    xp, itp = s.mass_solve(_lib.PHW_P, bp)
    xq, itq = s.mass_solve(_lib.PHW_Q, bq)
    assert rel(xp, spla.spsolve(ops['M_p'].tocsc(), bp)) < 1e-11
    assert rel(xq, spla.spsolve(ops['M_q'].tocsc(), bq)) < 1e-11
    assert 0 < itp < 80 and 0 < itq < 120, (itp, itq)
    # zero right-hand side converges immediately to zero
    z, it0 = s.mass_solve(_lib.PHW_Q, np.zeros(topo.n_edges))
    assert np.all(z == 0.0)


CASES = [
    # name, Nx, Ny, steps, tFinal, kwargs
    ('SE_wave', 34, 8, 120, 0.06, dict(timeIntScheme='SE')),
    ('EE_wave', 34, 8, 120, 0.06, dict(timeIntScheme='EE')),
    ('EH_wave', 34, 8, 120, 0.06, dict(timeIntScheme='EH')),
    ('SV_wave', 34, 8, 120, 0.06, dict(timeIntScheme='SV')),
    ('IE_wave', 34, 8, 120, 0.06, dict(timeIntScheme='IE')),
    ('SM_wave', 34, 8, 120, 0.06, dict(timeIntScheme='SM')),
    ('SV_strong', 34, 8, 120, 0.06, dict(timeIntScheme='SV', dirichletImp='strong')),
    ('SE_strong', 34, 8, 120, 0.06, dict(timeIntScheme='SE', dirichletImp='strong')),
    ('EH_strong', 34, 8, 120, 0.06, dict(timeIntScheme='EH', dirichletImp='strong')),
    ('SM_strong', 34, 8, 120, 0.06, dict(timeIntScheme='SM', dirichletImp='strong')),
    ('IE_strong', 34, 8, 120, 0.06, dict(timeIntScheme='IE', dirichletImp='strong')),
    ('SV_strong_analytical', 40, 10, 60, 0.03, dict(timeIntScheme='SV', dirichletImp='strong', analytical=True)),
    ('SE_IC', 52, 13, 100, 0.1, dict(timeIntScheme='SE', interConnection='IC')),
    ('SV_IC', 52, 13, 100, 0.1, dict(timeIntScheme='SV', interConnection='IC')),
    ('SM_IC', 52, 13, 100, 0.1, dict(timeIntScheme='SM', interConnection='IC')),
    ('SM_casimir', 40, 10, 100, 0.1, dict(timeIntScheme='SM', controlType='casimir')),
    ('SM_passivity', 40, 10, 150, 0.15, dict(timeIntScheme='SM', interConnection='IC', controlType='passivity',
                                             input_stop_t=0.02, control_start_t=0.06)),
    ('SE_passivity', 40, 10, 150, 0.15, dict(timeIntScheme='SE', interConnection='IC', controlType='passivity',
                                             input_stop_t=0.02, control_start_t=0.06)),
    ('SE_analytical', 40, 10, 60, 0.03, dict(timeIntScheme='SE', analytical=True)),
    ('SV_analytical', 40, 10, 60, 0.03, dict(timeIntScheme='SV', analytical=True)),
]


@pytest.mark.parametrize('name,Nx,Ny,steps,tF,kw', CASES)
def test_time_loop_parity(name, Nx, Ny, steps, tF, kw):
    mesh = rectangle_structured(Nx, Ny, 1.0, 0.25)
    Ho, nco, so = wave_2D_solve_oracle(tF, steps, mesh, 1.0, 0.25, K_wave=K_WAVE, rho=RHO, return_state=True, **kw)
    Hg, ncg, sg = wave_2D_solve(tF, steps, '/tmp', 0, 1.0, 0.25, K_wave=K_WAVE, rho=RHO, mesh=mesh, patch_cells=128,
                                return_state=True, verbose=False, **kw)
    assert ncg == nco
    assert Hg.shape == Ho.shape
    assert np.array_equal(Hg[:, 0], Ho[:, 0])                      # t accumulates identically
    assert rel(sg['p'], so['p']) < TOL_STATE, rel(sg['p'], so['p'])
    assert rel(sg['q'], so['q']) < TOL_STATE, rel(sg['q'], so['q'])
    if kw.get('interConnection') == 'IC':
        assert rel(sg['x'], so['x']) < TOL_STATE
    scale = np.abs(Ho[:, [1, 4, 5, 6, 7]]).max() + 1e-300
    for col in range(1, Ho.shape[1]):
        assert np.abs(Hg[:, col] - Ho[:, col]).max() < 1e-10 * max(scale, np.abs(Ho[:, col]).max()), (name, col)
    assert sg['stats']['max_rel_residual'] <= 1e-13


def test_time_loop_parity_unstructured_and_s1c():
    mesh = rectangle_delaunay(60, 15, seed=5)
    kw = dict(timeIntScheme='SV', interConnection='IC')
    Ho, _, so = wave_2D_solve_oracle(0.05, 50, mesh, 1.0, 0.25, K_wave=K_WAVE, rho=RHO, return_state=True, **kw)
    Hg, _, sg = wave_2D_solve(0.05, 50, '/tmp', 0, 1.0, 0.25, K_wave=K_WAVE, rho=RHO, mesh=mesh, patch_cells=200,
                              return_state=True, verbose=False, **kw)
    assert rel(sg['p'], so['p']) < TOL_STATE and rel(sg['q'], so['q']) < TOL_STATE and rel(sg['x'], so['x']) < TOL_STATE
    mesh = square_one_channel(20)
    Ho, _, so = wave_2D_solve_oracle(0.05, 50, mesh, 1.0, 0.1, domainShape='S_1C', K_wave=K_WAVE, rho=RHO, return_state=True, **kw)
    Hg, _, sg = wave_2D_solve(0.05, 50, '/tmp', 0, 1.0, 0.1, domainShape='S_1C', K_wave=K_WAVE, rho=RHO, mesh=mesh,
                              patch_cells=256, return_state=True, verbose=False, **kw)
    assert rel(sg['p'], so['p']) < TOL_STATE and rel(sg['q'], so['q']) < TOL_STATE and rel(sg['x'], so['x']) < TOL_STATE
    assert np.abs(Hg[:, 1:] - Ho[:, 1:]).max() < 1e-10 * np.abs(Ho[:, 1:]).max()


def test_sv_first_same_as_last_is_exact():
    """Stormer-Verlet reuses the last half-step velocity as the next first half-step (identical operands):
    switching the reuse off must give the same trajectory to solver tolerance."""
    mesh = rectangle_structured(34, 8)
    m = Mesh(*mesh, patch_cells=128)
    m.set_boundary(tag_boundary_edges(m, 'R', 1.0, 0.25))
    out = []
    for flags in (0, 1):
        prm = _lib.PhwParams()
        prm.scheme, prm.dt, prm.K_wave, prm.rho = _lib.SCHEMES['SV'], 5e-4, K_WAVE, RHO
        prm.input_kind, prm.in_amp, prm.in_omega, prm.in_t_stop = _lib.IN_SINE_WINDOW, 10.0, 8 * np.pi, 0.25
        prm.tol, prm.max_iter, prm.flags = 1e-13, 200, flags
        s = Solver(m, prm)
        H = s.run(80)
        out.append((H, s.get_state(), s.stats()))
    assert rel(out[0][1][0], out[1][1][0]) < 1e-11 and rel(out[0][1][1], out[1][1][1]) < 1e-11
    assert out[0][2]['solves_q'] < out[1][2]['solves_q']


def test_golden_exact_entries_on_gpu():
    """Mesh-independent entries of the reference's golden files (row 0 of the IC/SV traces)."""
    mesh = rectangle_structured(16, 4)
    H, _ = wave_2D_solve(0.1, 100, '/tmp', 0, 1.0, 0.25, K_wave=K_WAVE, rho=RHO, mesh=mesh, timeIntScheme='SV',
                         interConnection='IC', verbose=False)
    assert abs(H[0, 3] + 1.046651205470e-08) < 2e-20
    assert abs(H[0, 5] - 7.890733406804e-08) < 2e-19
    assert abs(H[0, 6] - 7.890733434191e-08) < 2e-19


def test_batched_sweep_matches_per_case_oracle():
    """BASELINE configs[4]: independent cases with different wave speed and lumped constants in one batched run."""
    from porthamiltonian_fem_b200.sweep import wave_sweep_solve, sweep_cases, LUMPED_KEYS, DEFAULT_LUMPED
    mesh = rectangle_structured(26, 6, 1.0, 0.25)
    cases = sweep_cases(7)
    cases[3] = dict(K_wave=3.0, rho=2.0)                       # the paper constants among them
    H, nc, solver = wave_sweep_solve(0.06, 60, 0, 1.0, 0.25, cases, mesh=mesh, patch_cells=96, return_solver=True)
    xs = solver.group_states()
    assert H.shape == (7, 60, 8)
    for i, c in enumerate(cases):
        lum = {k: c.get(k, DEFAULT_LUMPED[k]) for k in LUMPED_KEYS}
        Ho, _, so = wave_2D_solve_oracle(0.06, 60, mesh, 1.0, 0.25, timeIntScheme='SV', interConnection='IC',
                                         K_wave=c['K_wave'], rho=c['rho'], lumped=lum, return_state=True)
        scale = np.abs(Ho[:, 1:]).max()
        assert np.abs(H[i][:, 1:] - Ho[:, 1:]).max() < 1e-10 * scale, i
        assert np.array_equal(H[i][:, 0], Ho[:, 0])
        assert rel(xs[i], so['x']) < TOL_STATE
    solver.close()


def test_mass_solve_port_rhs_on_square_cells():
    """Regression: on square cells the port vector b_L is an eigenvector of the Jacobi-scaled RT1 mass matrix at the
    centre of the Chebyshev interval, so the first iterate is exact while the second is not; the accepted iterate must be
    the one whose residual was verified."""
    for Nx, Ny, pc in [(68, 17, 128), (136, 34, 2048)]:
        v, c = rectangle_structured(Nx, Ny)
        m, s = make_solver(v, c, 'R', 1.0, 0.25, pc)
        topo = fem.Topology(v, c)
        ops = fem.assemble_wave_operators(topo, fem.tag_boundary(topo, 'R', 1.0, 0.25))
        x, it = s.mass_solve(_lib.PHW_Q, ops['b_L'])
        assert rel(x, spla.spsolve(ops['M_q'].tocsc(), ops['b_L'])) < 1e-11, (Nx, Ny, it)


def _sliver_mesh():
    """a quality mesh with a band of stretched cells (aspect ratio ~25): three mesh columns squeezed to 4 % of their width"""
    v, c = rectangle_structured(40, 10)
    v = v.copy()
    x = v[:, 0]
    h = 1.0 / 40
    band = (x > 0.5 - 1e-9) & (x < 0.5 + 3 * h + 1e-9)
    v[band, 0] = 0.5 + (x[band] - 0.5) * 0.04
    right = x > 0.5 + 3 * h + 1e-9
    v[right, 0] = 0.5 + 3 * h * 0.04 + (x[right] - (0.5 + 3 * h)) * (0.5 - 3 * h * 0.04) / (0.5 - 3 * h)
    return v, c


def test_conjugate_gradient_mass_solves_on_poor_meshes():
    """the Chebyshev interval of the RT1 mass matrix is an element-wise bound, so a band of slivers (or uniformly
    stretched cells) makes every solve of the whole mesh slow; the reference's LU does not care (wave_2D.py:820,854).
    phw_solver_create switches the M_q solve to conjugate gradients (solve_pcg_coop_k) when the bound of kappa exceeds 10."""
    for name, (v, c) in (('aspect12', rectangle_structured(6, 18)), ('sliver_band', _sliver_mesh())):
        topo = fem.Topology(v, c)
        ops = fem.assemble_wave_operators(topo, fem.tag_boundary(topo, 'R', 1.0, 0.25))
        rng = np.random.default_rng(5)
        bp, bq = rng.standard_normal(topo.n_vertices), rng.standard_normal(topo.n_edges)
        ref_p, ref_q = spla.spsolve(ops['M_p'].tocsc(), bp), spla.spsolve(ops['M_q'].tocsc(), bq)
        its = {}
        for label, flags in (('auto', 0), ('chebyshev', 131072), ('cg_both', 65536)):
            m, s = make_solver(v, c, 'R', 1.0, 0.25, 64, flags=flags, max_iter=2000)
            st = s.stats()
            assert st['lam_max_q'] > 10 * st['lam_min_q'], name             # the element-wise bound flags the mesh
            xq, itq = s.mass_solve(_lib.PHW_Q, bq)
            xp, itp = s.mass_solve(_lib.PHW_P, bp)
            assert rel(xq, ref_q) < 1e-10 and rel(xp, ref_p) < 1e-11, (name, label, rel(xq, ref_q), rel(xp, ref_p))
            z, _ = s.mass_solve(_lib.PHW_Q, np.zeros(topo.n_edges))
            assert np.all(z == 0.0)
            its[label] = (itp, itq, bool(s.stats()['kernel_paths'] & 2048))
            s.close(); m.close()
        assert its['auto'][2] and its['cg_both'][2] and not its['chebyshev'][2], its      # PHW_PATH_PCG
        assert its['auto'][1] == its['cg_both'][1] and its['auto'][0] == its['chebyshev'][0], its   # auto: CG for M_q only
        # uniformly stretched cells: the Jacobi-scaled matrix IS ill-conditioned (true kappa 119 on this mesh), CG gains little;
        # a band of slivers only adds outlying eigenvalues, which CG deflates (host CG: 103 iterations) while the Chebyshev interval
        # is set by the worst cell
        assert its['auto'][1] <= its['chebyshev'][1], (name, its)
        if name == 'sliver_band':
            assert its['auto'][1] < 0.5 * its['chebyshev'][1], its


def test_time_loop_on_sliver_mesh_matches_oracle_and_warns():
    mesh = _sliver_mesh()
    steps, tF = 60, 0.03
    for scheme, kw in (('SE', {}), ('SV', dict(interConnection='IC'))):
        Ho, _, so = wave_2D_solve_oracle(tF, steps, mesh, 1.0, 0.25, K_wave=K_WAVE, rho=RHO, timeIntScheme=scheme, return_state=True, **kw)
        with pytest.warns(UserWarning, match='conjugate'):
            Hg, _, sg = wave_2D_solve(tF, steps, '/tmp', 0, 1.0, 0.25, K_wave=K_WAVE, rho=RHO, mesh=mesh, timeIntScheme=scheme,
                                      return_state=True, verbose=False, **kw)
        assert sg['stats']['kernel_paths'] & 2048
        assert rel(sg['p'], so['p']) < 1e-9 and rel(sg['q'], so['q']) < 1e-9, (rel(sg['p'], so['p']), rel(sg['q'], so['q']))
        assert np.abs(Hg[:, 1:] - Ho[:, 1:]).max() < 1e-9 * np.abs(Ho[:, 1:]).max()


@pytest.mark.parametrize('scheme', ['SE', 'SV', 'EE'])
def test_time_loop_parity_config1_size(scheme):
    mesh = rectangle_structured(136, 34, 1.0, 0.25)
    steps, tF = 150, 0.075
    Ho, _, so = wave_2D_solve_oracle(tF, steps, mesh, 1.0, 0.25, K_wave=K_WAVE, rho=RHO, timeIntScheme=scheme, return_state=True)
    Hg, _, sg = wave_2D_solve(tF, steps, '/tmp', 0, 1.0, 0.25, K_wave=K_WAVE, rho=RHO, mesh=mesh, timeIntScheme=scheme,
                              return_state=True, verbose=False)
    assert rel(sg['p'], so['p']) < TOL_STATE and rel(sg['q'], so['q']) < TOL_STATE
    assert np.abs(Hg[:, 1:] - Ho[:, 1:]).max() < 1e-10 * np.abs(Ho[:, 1:]).max()


def test_edge_cases_empty_run_tiny_meshes_and_restart():
    """n_steps = 0, a two-cell mesh, a one-patch mesh with the smallest patch size, and restart through
    phw_get_state / phw_set_state / phw_set_time: two half runs reproduce one full run."""
    v, c = rectangle_structured(1, 1)
    m, s = make_solver(v, c, 'R', 1.0, 0.25, 16, scheme='SV')
    assert s.run(0).shape == (0, 8)
    H = s.run(3)
    Ho, _ = wave_2D_solve_oracle(3e-3, 3, (v, c), 1.0, 0.25, K_wave=K_WAVE, rho=RHO, timeIntScheme='SV')
    assert np.abs(H[:, 1:] - Ho[:, 1:]).max() <= 1e-10 * max(np.abs(Ho[:, 1:]).max(), 1e-300)
    # restart: 40 steps == 20 + (state, clock, accumulators carried over by the caller) + 20
    v, c = rectangle_delaunay(30, 8, seed=11)
    m, s = make_solver(v, c, 'R', 1.0, 0.25, 80, scheme='SE')
    H_full = s.run(40)
    p_full, q_full, _ = s.get_state()
    m2, s2 = make_solver(v, c, 'R', 1.0, 0.25, 80, scheme='SE')
    H_a = s2.run(20)
    p, q, x = s2.get_state()
    m3, s3 = make_solver(v, c, 'R', 1.0, 0.25, 80, scheme='SE')
    s3.set_state(p=p, q=q, x3=x)
    s3.set_time(H_a[-1, 0])
    H_b = s3.run(20)
    p_b, q_b, _ = s3.get_state()
    assert rel(p_b, p_full) < 1e-10 and rel(q_b, q_full) < 1e-10
    assert np.allclose(H_b[:, 0], H_full[20:, 0], rtol=0, atol=1e-15)
    assert np.abs(H_b[:, 7] - H_full[20:, 7]).max() < 1e-10 * np.abs(H_full[:, 7]).max()      # H_wave depends on the state only
