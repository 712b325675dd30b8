"""Higher-order pairs P_k x RT_k (basis_order of wave_2D.py:15-20,174-175), CPU side:
  * the oracle's restatement (oracle/fem_ho.py: reference element + Piola map) pinned at discretisation level against
    the reference's golden traces that use a higher-order pair (run_test_cases.py:30-31);
  * the product's host assembly (porthamiltonian_fem_b200/fem_ho.py: physical-cell construction) against the oracle's,
    operator by operator, including the integer dof maps (bit-exact)."""
import os

import numpy as np
import pytest

from oracle import fem, fem_ho as ofem
from oracle.wave_2D_oracle import wave_2D_solve_oracle
from porthamiltonian_fem_b200.fem_ho import HighOrderSpaces
from porthamiltonian_fem_b200.mesh import rectangle_structured, rectangle_delaunay

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'reference_H_arrays.npz'))
PAIRS = [(1, 1), (2, 2), (3, 3), (1, 2), (2, 1), (3, 2)]


@pytest.mark.parametrize('kp,kq', PAIRS)
@pytest.mark.parametrize('meshname', ['structured', 'delaunay'])
def test_host_assembly_matches_oracle(kp, kq, meshname):
    mesh = rectangle_structured(8, 2, 1.0, 0.25) if meshname == 'structured' else rectangle_delaunay(10, 3, seed=3)
    topo = fem.Topology(*mesh)
    tags = fem.tag_boundary(topo, 'R', 1.0, 0.25)
    so = ofem.SpacesHO(topo, kp, kq)
    S = HighOrderSpaces(mesh[0], mesh[1], kp, kq)
    assert np.array_equal(S.edges, topo.edges) and np.array_equal(S.boundary_edges, topo.boundary_edges)
    assert (S.np_dofs, S.nq_dofs) == (so.n_p, so.n_q)
    # same set of global dofs per cell (local orderings differ by construction)
    assert np.array_equal(np.sort(S.cell_dofs_p, axis=1), np.sort(so.cell_dofs_p, axis=1))
    assert np.array_equal(np.sort(S.cell_dofs_q, axis=1), np.sort(so.cell_dofs_q, axis=1))
    assert np.abs(S.p_coords - so.p_coords).max() < 1e-15
    w = lambda x, y: 2.0 + np.sin(9.0 * y + 0.3)
    oo = so.assemble(tags, left_weight=w)
    po = S.assemble({int(e): int(t) for e, t in zip(S.boundary_edges, tags)}, left_weight=w)
    for k in ('M_p', 'M_q', 'D'):
        assert abs(oo[k] - po[k]).max() < 1e-12 * abs(oo[k]).max(), k
    assert np.abs(oo['b_L'] - po['b_L']).max() < 1e-12 * np.abs(oo['b_L']).max()
    # Wathen bounds enclose the Jacobi-scaled spectra
    (lp, hp), (lq, hq) = S.jacobi_bounds(po)
    for M, lo, hi in ((oo['M_p'], lp, hp), (oo['M_q'], lq, hq)):
        d = 1.0 / np.sqrt(M.diagonal())
        ev = np.linalg.eigvalsh((M.multiply(d[:, None]).multiply(d[None, :])).toarray())
        assert lo * (1 - 1e-10) <= ev.min() and ev.max() <= hi * (1 + 1e-10)


def test_lowest_order_reduces_to_p1_rt1_oracle():
    mesh = rectangle_delaunay(12, 4, seed=1)
    topo = fem.Topology(*mesh)
    tags = fem.tag_boundary(topo, 'R', 1.0, 0.25)
    a = ofem.SpacesHO(topo, 1, 1).assemble(tags)
    b = fem.assemble_wave_operators(topo, tags)
    for k in ('M_p', 'M_q', 'D'):
        assert abs(a[k] - b[k]).max() < 1e-13 * abs(b[k]).max()
    assert np.abs(a['b_L'] - b['b_L']).max() < 1e-14


def test_discrete_energy_balance_high_order():
    """SV with exact solves: the wave-only energy residual stays O(dt^2) for any conforming pair."""
    H, _ = wave_2D_solve_oracle(0.05, 100, rectangle_structured(8, 2, 1.0, 0.25), 1.0, 0.25, K_wave=3.0, rho=2.0,
                                timeIntScheme='SV', basis_order=(2, 2))
    assert np.abs(H[:, 2]).max() < 1e-3 * np.abs(H[:, 1]).max()


@pytest.mark.parametrize('case,bo,Nx,Ny', [
    ('R_SV_weak_nx4_analytical_t0_1_steps200_P3_RT3', (3, 3), 8, 2),
    ('R_SV_weak_nx20_analytical_t0_1_steps200_P1_RT2', (1, 2), 34, 8),
])
def test_golden_high_order_discretisation_level(case, bo, Nx, Ny):
    Hg = G['test_cases/' + case + '/H']
    H, _ = wave_2D_solve_oracle(0.1, 200, rectangle_structured(Nx, Ny, 1.0, 0.25), 1.0, 0.25, K_wave=3.0, rho=2.0,
                                timeIntScheme='SV', analytical=True, basis_order=bo)
    assert np.array_equal(H[:, 0], Hg[:, 0])
    assert np.abs(H[:, 1] - Hg[:, 1]).max() < 0.01 * Hg[:, 1].max()               # Hamiltonian trace
    assert abs(H[0, 2] - Hg[0, 2]) < 0.1 * abs(Hg[0, 2])                           # first-step energy residual (sign + size)
    assert abs(H[0, 8] - Hg[0, 8]) < 0.1 * Hg[0, 8]                                # initial L2 interpolation error
    assert abs(H[0, 11] - Hg[0, 11]) < 0.1 * abs(Hg[0, 11])                        # exact solution at the probe point


def test_convergence_order_of_the_interpolant():
    """L2 interpolation error of the analytical initial state falls like h^(k+1) (plotting_wave_2D.py:403-415 trend lines)."""
    for k, rate in ((1, 2.0), (2, 3.0), (3, 4.0)):
        errs = []
        for Nx, Ny in ((16, 4), (32, 8)):
            H, _ = wave_2D_solve_oracle(1e-6, 1, rectangle_structured(Nx, Ny, 1.0, 0.25), 1.0, 0.25, K_wave=3.0, rho=2.0,
                                        timeIntScheme='SE', analytical=True, basis_order=(k, k))
            errs.append(H[0, 8])
        assert abs(np.log2(errs[0] / errs[1]) - rate) < 0.35, (k, errs)
