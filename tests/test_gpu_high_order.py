"""GPU parity of the higher-order pairs P_k x RT_k (basis_order, wave_2D.py:15-20,174-175): the device
CSR path (phw_ho_*) against the CPU oracle (oracle/fem_ho.py, independent assembly + sparse LU of the mixed system)
on the same meshes: state relative L2 <= 1e-10, every H_array column (incl. the analytical diagnostics) to 1e-10."""
import numpy as np
import pytest

from oracle.wave_2D_oracle import wave_2D_solve_oracle
from porthamiltonian_fem_b200 import wave_2D_solve
from porthamiltonian_fem_b200.mesh import rectangle_structured, rectangle_delaunay

pytestmark = pytest.mark.gpu
K_WAVE, RHO = 3.0, 2.0


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


CASES = [
    ('SV_analytical_P3RT3_nx4', (3, 3), lambda: rectangle_structured(8, 2), 200, 0.1, dict(timeIntScheme='SV', analytical=True)),
    ('SV_analytical_P1RT2_nx20', (1, 2), lambda: rectangle_structured(34, 8), 100, 0.05, dict(timeIntScheme='SV', analytical=True)),
    ('SE_analytical_P3RT3', (3, 3), lambda: rectangle_structured(16, 4), 100, 0.02, dict(timeIntScheme='SE', analytical=True)),
    ('SE_analytical_P2RT2', (2, 2), lambda: rectangle_structured(20, 5), 100, 0.03, dict(timeIntScheme='SE', analytical=True)),
    ('SV_wave_P2RT1', (2, 1), lambda: rectangle_structured(20, 5), 100, 0.03, dict(timeIntScheme='SV')),
    ('EH_wave_P2RT2_delaunay', (2, 2), lambda: rectangle_delaunay(24, 6, seed=2), 100, 0.03, dict(timeIntScheme='EH')),
    ('EE_wave_P3RT2', (3, 2), lambda: rectangle_structured(12, 3), 80, 0.02, dict(timeIntScheme='EE')),
    ('SV_IC_P2RT2', (2, 2), lambda: rectangle_structured(20, 5), 100, 0.1, dict(timeIntScheme='SV', interConnection='IC')),
    ('SE_IC_P3RT3', (3, 3), lambda: rectangle_structured(12, 3), 100, 0.05, dict(timeIntScheme='SE', interConnection='IC')),
]


@pytest.mark.parametrize('name,bo,make,steps,tF,kw', CASES)
def test_high_order_time_loop_parity(name, bo, make, steps, tF, kw, monkeypatch):
    monkeypatch.delenv('PHW_HO_CSR', raising=False)
    monkeypatch.setenv('PHW_HO_MF', '1')                     # the matrix-free kernels (csrc/kernels_ho.cuh) at every order
    _parity(name, bo, make, steps, tF, kw, matrix_free=True)


@pytest.mark.parametrize('name,bo,make,steps,tF,kw', [CASES[0], CASES[3], CASES[4]])
def test_high_order_default_path_selection(name, bo, make, steps, tF, kw, monkeypatch):
    """default: matrix-free when an order is 3, assembled rows below (measured: profiles/r2o_ho_cases_matrix_free_vs_csr.jsonl)"""
    monkeypatch.delenv('PHW_HO_CSR', raising=False)
    monkeypatch.delenv('PHW_HO_MF', raising=False)
    _parity(name, bo, make, steps, tF, kw, matrix_free=max(bo) >= 3)


@pytest.mark.parametrize('name,bo,make,steps,tF,kw', [CASES[0], CASES[5], CASES[8]])
def test_high_order_time_loop_parity_assembled_csr(name, bo, make, steps, tF, kw, monkeypatch):
    monkeypatch.delenv('PHW_HO_MF', raising=False)
    monkeypatch.setenv('PHW_HO_CSR', '1')                    # host-assembled CSR operators
    _parity(name, bo, make, steps, tF, kw, matrix_free=False)


IMPLICIT_CASES = [
    ('SM_wave_P2RT2', (2, 2), lambda: rectangle_structured(20, 5), 60, 0.03, dict(timeIntScheme='SM')),
    ('IE_wave_P3RT2_delaunay', (3, 2), lambda: rectangle_delaunay(16, 4, seed=5), 60, 0.02, dict(timeIntScheme='IE')),
    ('SM_analytical_P3RT3', (3, 3), lambda: rectangle_structured(12, 3), 60, 0.02, dict(timeIntScheme='SM', analytical=True)),
    ('SM_IC_P2RT2', (2, 2), lambda: rectangle_structured(20, 5), 80, 0.08, dict(timeIntScheme='SM', interConnection='IC')),
    ('SM_IC_P3RT3', (3, 3), lambda: rectangle_structured(8, 2), 60, 0.05, dict(timeIntScheme='SM', interConnection='IC')),
]


@pytest.mark.parametrize('name,bo,make,steps,tF,kw', IMPLICIT_CASES)
def test_high_order_implicit_schemes_parity(name, bo, make, steps, tF, kw, monkeypatch):
    """IE / SM (wave_2D.py:305-319,560-572) at basis_order > 1: coupled block solve on the assembled CSR rows (solve_block_csr_k),
    with the bordered lumped rows for 'IC'"""
    monkeypatch.delenv('PHW_HO_CSR', raising=False)
    monkeypatch.setenv('PHW_HO_MF', '1')                     # even when the matrix-free path is asked for
    _parity(name, bo, make, steps, tF, kw, matrix_free=False, tol=2e-10)


def _parity(name, bo, make, steps, tF, kw, matrix_free, tol=1e-10):
    mesh = make()
    Ho, nco, so = wave_2D_solve_oracle(tF, steps, mesh, 1.0, 0.25, K_wave=K_WAVE, rho=RHO, basis_order=bo, return_state=True, **kw)
    Hg, ncg, sg = wave_2D_solve(tF, steps, '/tmp', 0, 1.0, 0.25, K_wave=K_WAVE, rho=RHO, mesh=mesh, basis_order=bo,
                                return_state=True, verbose=False, **kw)
    assert ncg == nco and Hg.shape == Ho.shape
    assert np.array_equal(Hg[:, 0], Ho[:, 0])
    assert rel(sg['p'], so['p']) < tol, rel(sg['p'], so['p'])
    assert rel(sg['q'], so['q']) < tol, rel(sg['q'], so['q'])
    if kw.get('interConnection') == 'IC':
        assert rel(sg['x'], so['x']) < tol
    scale = np.abs(Ho[:, [1, 4, 5, 6, 7]]).max() + 1e-300
    for col in range(1, Ho.shape[1]):
        assert np.abs(Hg[:, col] - Ho[:, col]).max() < tol * max(scale, np.abs(Ho[:, col]).max()), (name, col)
    assert sg['stats']['max_rel_residual'] <= 1e-13
    assert bool(sg['stats']['kernel_paths'] & 1024) == matrix_free      # PHW_PATH_HO_MF


def test_high_order_rejects_unimplemented_variants():
    mesh = rectangle_structured(4, 1)
    for kw in (dict(timeIntScheme='SM', interConnection='IC', controlType='passivity', input_stop_t=0.0, control_start_t=1.0),
               dict(timeIntScheme='SV', dirichletImp='strong')):
        with pytest.raises(NotImplementedError):
            wave_2D_solve(0.01, 2, '/tmp', 0, 1.0, 0.25, mesh=mesh, basis_order=(2, 2), verbose=False, **kw)


@pytest.mark.parametrize('bo', [(2, 2), (3, 3), (1, 2)])
def test_matrix_free_operators_match_reference_assembly(bo):
    """kernel-level entry points on the matrix-free solver: D, D^T, M_p, M_q applied by the cell kernels against the same operators
    assembled on the host from the reference matrices (fem_ref.RefSpaces.assemble), and the Chebyshev mass solves against sparse LU"""
    import scipy.sparse.linalg as spla
    from porthamiltonian_fem_b200 import _lib
    from porthamiltonian_fem_b200.fem_ref import RefSpaces
    from porthamiltonian_fem_b200.solver import MatrixFreeHighOrderSolver
    try:
        from test_fem_ref_cpu import tags_of
    except ImportError:
        from tests.test_fem_ref_cpu import tags_of
    mesh = rectangle_delaunay(12, 4, seed=4)
    S = RefSpaces(mesh[0], mesh[1], bo[0], bo[1])
    bops = S.boundary_operators(tags_of(S))
    ops = S.assemble(bops)
    prm = _lib.PhwParams()
    prm.scheme, prm.dt, prm.K_wave, prm.rho, prm.tol, prm.max_iter = _lib.SCHEMES['SE'], 1e-3, 1.0, 1.0, 1e-13, 2000
    bp, bq = S.jacobi_bounds()
    sol = MatrixFreeHighOrderSolver(S, bops, prm, bp, bq)
    rng = np.random.default_rng(1)
    p, q = rng.standard_normal(S.np_dofs), rng.standard_normal(S.nq_dofs)
    assert rel(sol.apply_D(p), ops['D'] @ p) < 1e-13
    assert rel(sol.apply_DT(q), ops['D'].T @ q) < 1e-13
    assert rel(sol.mass_matvec(_lib.PHW_P, p), ops['M_p'] @ p) < 1e-13
    assert rel(sol.mass_matvec(_lib.PHW_Q, q), ops['M_q'] @ q) < 1e-13
    xp, itp = sol.mass_solve(_lib.PHW_P, p)
    xq, itq = sol.mass_solve(_lib.PHW_Q, q)
    assert rel(xp, spla.spsolve(ops['M_p'].tocsc(), p)) < 1e-11 and rel(xq, spla.spsolve(ops['M_q'].tocsc(), q)) < 1e-11
    assert 0 < itp < 200 and 0 < itq < 300
    sol.close()
