"""Eigenvalue check on the device operators (reference calculate_eigen.py:153-228, main_eigen.py): the spectrum of the
semi-discrete wave operator built from phw_apply_D / phw_apply_DT / phw_mass_solve against (i) the same pencil assembled by the CPU
oracle and (ii) the exact eigenvalues c sqrt((j pi/Lx)^2 + (i pi/Ly)^2)."""
import numpy as np
import pytest
import scipy.linalg as sla

from oracle import fem
from porthamiltonian_fem_b200.calculate_eigen import calculate_eigen, exact_eigenvalues
from porthamiltonian_fem_b200.mesh import rectangle_structured

pytestmark = pytest.mark.gpu
K_WAVE, RHO = 3.0, 2.0


def oracle_omegas(mesh, n):
    topo = fem.Topology(*mesh)
    ops = fem.assemble_wave_operators(topo, fem.tag_boundary(topo, 'R', 1.0, 0.25))
    Mq, Mp, D = ops['M_q'].toarray(), ops['M_p'].toarray(), ops['D'].toarray()
    S = D.T @ np.linalg.solve(Mq, D)
    lam = sla.eigh(0.5 * (S + S.T), Mp, eigvals_only=True)
    lam = lam[lam > 1e-8]
    return np.sqrt(K_WAVE / RHO) * np.sqrt(lam[:n])


def fem_ho_omegas(fem_ho, mesh, n, k=2):
    """the same pencil for P_k x RT_k, assembled by the CPU oracle (oracle/fem_ho.py) with the facet markers of the time loop"""
    topo = fem.Topology(*mesh)
    ops = fem_ho.SpacesHO(topo, k, k).assemble(fem.tag_boundary(topo, 'R', 1.0, 0.25))
    Mq, Mp, D = ops['M_q'].toarray(), ops['M_p'].toarray(), ops['D'].toarray()
    S = D.T @ np.linalg.solve(Mq, D)
    lam = sla.eigh(0.5 * (S + S.T), Mp, eigvals_only=True)
    lam = lam[lam > 1e-8]
    return np.sqrt(K_WAVE / RHO) * np.sqrt(lam[:n])


@pytest.mark.parametrize('method', ['dense', 'lobpcg'])
def test_lowest_order_spectrum(method):
    mesh = rectangle_structured(24, 6)
    out, nc = calculate_eigen(0.5, 1000, None, 0, 1.0, 0.25, K_wave=K_WAVE, rho=RHO, mesh=mesh, num_eigs=12, method=method)
    assert nc == len(mesh[1]) and out.shape == (12, 3)
    ref = oracle_omegas(mesh, 12)
    assert np.abs(out[:, 1] - ref).max() < (1e-8 if method == 'dense' else 1e-5) * ref.max()
    assert np.allclose(out[:, 0], exact_eigenvalues(np.sqrt(K_WAVE / RHO), 1.0, 0.25, 12))
    # discretisation level on this mesh: every one of the first exact eigenvalues has a model eigenvalue next to it (the P1 x RT1
    # pencil also carries a second branch close to some of them, so the sorted lists are not compared entry by entry)
    for ex in out[:4, 0]:
        assert np.abs(out[:, 1] - ex).min() < 0.02 * ex
    assert np.abs(out[:, 2]).max() < 1e-6 * out[:, 1].max()                  # purely imaginary pairs


def test_second_order_pair_as_in_main_eigen():
    """main_eigen.py runs P2 x RT2 on nx = 5 .. 25.  The device spectrum equals the one of the pencil assembled on the CPU (same
    boundary tags), and the error of the lowest modes drops with the mesh width (4.5 .. 5.5 x between these two pre-asymptotic meshes,
    ~13 x = h^4 between the next two: checked on the CPU pencil, where a 64 x 16 mesh is cheap)"""
    from oracle import fem_ho
    errs = []
    for n in (4, 8):
        mesh = rectangle_structured(4 * n, n)
        out, _ = calculate_eigen(0.5, 1000, None, 0, 1.0, 0.25, K_wave=K_WAVE, rho=RHO, mesh=mesh, basis_order=(2, 2), num_eigs=8, method='dense')
        errs.append(max(np.abs(out[:, 1] - ex).min() / ex for ex in out[:3, 0]))
        ref = fem_ho_omegas(fem_ho, mesh, 8)
        assert np.abs(out[:, 1] - ref).max() < 1e-8 * ref.max()
    assert errs[1] < 2e-6 and errs[0] / errs[1] > 4.0, errs
