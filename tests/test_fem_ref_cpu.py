"""Reference-element form of the higher-order pairs (porthamiltonian_fem_b200/fem_ref.py: what the matrix-free device kernels apply)
against the physical-cell construction of fem_ho.py on the same meshes: after the cell-local change of the interior RT dofs to the
dofs of the specification, M_p, M_q, D = C - N and b_L coincide to round-off; conversions round-trip; P dof coordinates coincide."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from porthamiltonian_fem_b200.fem_ho import HighOrderSpaces
from porthamiltonian_fem_b200.fem_ref import RefSpaces, RefElement
from porthamiltonian_fem_b200.mesh import rectangle_structured, rectangle_delaunay


def tags_of(S):
    be = S.boundary_edges
    pa, pb = S.vertices[S.edges[be, 0]], S.vertices[S.edges[be, 1]]
    tags = {}
    for e, a, b in zip(be, pa, pb):
        if abs(a[0]) < 1e-12 and abs(b[0]) < 1e-12:
            tags[int(e)] = 0
        elif abs(a[0] - 1) < 1e-12 and abs(b[0] - 1) < 1e-12:
            tags[int(e)] = 2
        elif abs(a[1] - 0.25) < 1e-12 and abs(b[1] - 0.25) < 1e-12:
            tags[int(e)] = 1
        else:
            tags[int(e)] = 3
    return tags


def rel(A, B):
    A = A.toarray() if hasattr(A, 'toarray') else np.asarray(A)
    B = B.toarray() if hasattr(B, 'toarray') else np.asarray(B)
    return np.abs(A - B).max() / np.abs(B).max()


@pytest.mark.parametrize('kp,kq,mesh', [(1, 1, 'structured'), (2, 2, 'delaunay'), (3, 3, 'structured'), (1, 2, 'structured'), (3, 2, 'delaunay'), (2, 3, 'delaunay')])
def test_reference_form_reproduces_the_assembled_operators(kp, kq, mesh):
    m = rectangle_structured(5, 2) if mesh == 'structured' else rectangle_delaunay(7, 3, seed=kp + kq)
    R = RefSpaces(m[0], m[1], kp, kq)
    H = HighOrderSpaces(m[0], m[1], kp, kq)
    assert np.array_equal(R.edges, H.edges) and R.np_dofs == H.np_dofs and R.nq_dofs == H.nq_dofs
    assert np.abs(R.p_coords - H.p_coords).max() < 1e-14
    lw = lambda x, y: 2.0 * np.sin(2 * np.pi * y / 0.25 + np.pi / 2)
    tags = tags_of(R)
    opsR = R.assemble(R.boundary_operators(tags, left_weight=lw))
    opsH = H.assemble(tags, left_weight=lw)
    T = R.spec_transform()                       # q_spec = T q_ref
    Ti = spla.inv(T.tocsc()).tocsr() if kq > 1 else T
    assert rel(opsR['M_p'], opsH['M_p']) < 1e-13
    assert rel(Ti.T @ opsR['M_q'] @ Ti, opsH['M_q']) < 1e-12
    assert rel(Ti.T @ opsR['D'], opsH['D']) < 1e-12
    assert np.abs(Ti.T @ opsR['b_L'] - opsH['b_L']).max() < 1e-13 and np.abs(Ti.T @ opsR['b_L1'] - opsH['b_L1']).max() < 1e-13
    q = np.random.default_rng(0).standard_normal(R.nq_dofs)
    assert np.abs(R.from_spec(R.to_spec(q)) - q).max() < 1e-12 and np.abs(T @ q - R.to_spec(q)).max() < 1e-12
    # spectral bounds of the Jacobi-scaled element matrices: valid interval, of the same order as in the physical-moment basis
    (lp0, lp1), (lq0, lq1) = R.jacobi_bounds()
    (_, _), (hq0, hq1) = H.jacobi_bounds(opsH)
    assert 0 < lp0 < lp1 and 0 < lq0 < lq1 and lq1 / lq0 < 4.0 * hq1 / hq0


def test_interpolation_and_normal_traces():
    """edge moments of an interpolated field are those of the field; a constant field is reproduced exactly (RT_k contains P_0^2)"""
    m = rectangle_delaunay(6, 3, seed=5)
    for k in (1, 2, 3):
        R = RefSpaces(m[0], m[1], 1, k)
        const = lambda x, y: (0.3 + 0 * x, -1.1 + 0 * x)
        q = R.interpolate_q(const)
        # evaluate the discrete field at the cell quadrature points
        r = R.ref
        phi_hat, _ = r.tab_q(r.q_bary[:, 1], r.q_bary[:, 2])
        phi = np.einsum('cij,qnj->cqni', R.J, phi_hat) / R.detJ[:, None, None, None]
        val = np.einsum('cqni,cn->cqi', phi, q[R.cell_dofs_q])
        assert np.abs(val[..., 0] - 0.3).max() < 1e-12 and np.abs(val[..., 1] + 1.1).max() < 1e-12


def test_reference_element_sizes():
    for kp, kq in [(1, 1), (2, 2), (3, 3)]:
        r = RefElement(kp, kq)
        assert r.ndp == (kp + 1) * (kp + 2) // 2 and r.ndq == kq * (kq + 2)
        assert abs(r.Mp_hat.sum() - 0.5) < 1e-14                       # partition of unity: area of the reference triangle
        assert np.allclose(r.C_hat.sum(axis=1), np.einsum('q,qj->j', r.q_w, r.tab_q(r.q_bary[:, 1], r.q_bary[:, 2])[1]))
