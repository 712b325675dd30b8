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
def test_high_order_time_loop_parity(name, bo, make, steps, tF, kw):
    mesh = make()
    Ho, nco, so = wave_2D_solve_oracle(tF, steps, mesh, 1.0, 0.25, K_wave=K_WAVE, rho=RHO, basis_order=bo, return_state=True, **kw)
    Hg, ncg, sg = wave_2D_solve(tF, steps, '/tmp', 0, 1.0, 0.25, K_wave=K_WAVE, rho=RHO, mesh=mesh, basis_order=bo,
                                return_state=True, verbose=False, **kw)
    assert ncg == nco and Hg.shape == Ho.shape
    assert np.array_equal(Hg[:, 0], Ho[:, 0])
    assert rel(sg['p'], so['p']) < 1e-10, rel(sg['p'], so['p'])
    assert rel(sg['q'], so['q']) < 1e-10, rel(sg['q'], so['q'])
    if kw.get('interConnection') == 'IC':
        assert rel(sg['x'], so['x']) < 1e-10
    scale = np.abs(Ho[:, [1, 4, 5, 6, 7]]).max() + 1e-300
    for col in range(1, Ho.shape[1]):
        assert np.abs(Hg[:, col] - Ho[:, col]).max() < 1e-10 * max(scale, np.abs(Ho[:, col]).max()), (name, col)
    assert sg['stats']['max_rel_residual'] <= 1e-13


def test_high_order_rejects_unimplemented_variants():
    mesh = rectangle_structured(4, 1)
    for kw in (dict(timeIntScheme='SM'), dict(timeIntScheme='SV', dirichletImp='strong')):
        with pytest.raises(NotImplementedError):
            wave_2D_solve(0.01, 2, '/tmp', 0, 1.0, 0.25, mesh=mesh, basis_order=(2, 2), verbose=False, **kw)
