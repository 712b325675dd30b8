"""Eigenvalue check of the semi-discrete wave operator (reference: calculate_eigen.py:14-228, driver main_eigen.py).

The reference assembles A and M of  M x_dot = A x  (x = (p, q); forms :171-174, no boundary data), forms the dense M^-1 A with
``np.linalg.inv`` and takes ``np.linalg.eig``: the spectrum is purely imaginary, +-i omega_j, and omega_j is compared with the exact
c sqrt((j pi / Lx)^2 + (i pi / Ly)^2), j >= 1, i >= 0 (:153-159).  Here the same operator is the one the time loop uses,

    M_p p_dot = -K D^T q,   M_q q_dot = (1/rho) D p      =>     omega^2 = (K/rho) lambda(M_p^-1 D^T M_q^-1 D),

and every factor is applied ON THE DEVICE through the kernel-level entry points of libphwave (phw_apply_D / phw_apply_DT: the
matrix-free skew operator; phw_mass_solve: the Chebyshev mass solves; higher-order pairs: the CSR kernels).  ``method='dense'``
builds the N_p x N_p matrix column by column and calls ``numpy.linalg.eig`` as the reference does (small meshes);
``method='lobpcg'`` finds the lowest modes of the symmetric pencil (D^T M_q^-1 D, M_p) with scipy's LOBPCG on device-backed operators.
"""
import numpy as np

from . import _lib
from .mesh import generate_mesh
from .solver import Mesh, Solver, tag_boundary_edges, tag_space_boundary


def exact_eigenvalues(c, xLength, yLength, n=None):
    """calculate_eigen.py:153-159"""
    vals = sorted(c * np.sqrt((JJ * np.pi / xLength) ** 2 + (II * np.pi / yLength) ** 2) for II in range(0, 15) for JJ in range(1, 40))
    return np.array(vals if n is None else vals[:n])


class _DeviceOperators:
    """x -> D x, D^T x, M^-1 x, M x on the device (P1/RT1: matrix-free kernels; P_k/RT_k: CSR kernels)"""

    def __init__(self, mesh, domainShape, xLength, yLength, basis_order, device):
        self.high_order = tuple(basis_order) != (1, 1)
        prm = _lib.PhwParams()
        prm.scheme, prm.dt, prm.K_wave, prm.rho = _lib.SCHEMES['SE'], 1e-3, 1.0, 1.0
        prm.tol, prm.max_iter, prm.extrap_order = 1e-13, 2000, 0
        if self.high_order:
            from .fem_ho import HighOrderSpaces
            from .solver import HighOrderSolver
            S = HighOrderSpaces(mesh[0], mesh[1], basis_order[0], basis_order[1])
            tags = tag_space_boundary(S, domainShape, xLength, yLength)
            ops = S.assemble(tags)
            bp, bq = S.jacobi_bounds(ops)
            self.solver = HighOrderSolver(S, ops, prm, bp, bq, device=device)
            self.mesh_handle = None
            self.n_p, self.n_q, self.n_cells = S.np_dofs, S.nq_dofs, S.nc
        else:
            self.mesh_handle = Mesh(mesh[0], mesh[1], patch_cells=max(64, min(2048, len(mesh[1]) // 16)))
            self.mesh_handle.set_boundary(tag_boundary_edges(self.mesh_handle, domainShape, xLength, yLength))
            self.solver = Solver(self.mesh_handle, prm, device=device)
            self.n_p, self.n_q, self.n_cells = self.mesh_handle.n_vertices, self.mesh_handle.n_edges, self.mesh_handle.n_cells

    def D(self, p):
        return self.solver.apply_D(np.ascontiguousarray(p, dtype=np.float64))

    def DT(self, q):
        return self.solver.apply_DT(np.ascontiguousarray(q, dtype=np.float64))

    def Minv(self, which, b):
        return self.solver.mass_solve(which, np.ascontiguousarray(b, dtype=np.float64))[0]

    def M(self, which, x):
        return self.solver.mass_matvec(which, np.ascontiguousarray(x, dtype=np.float64))

    def stiffness(self, p):
        """D^T M_q^-1 D p  (symmetric positive semi-definite)"""
        return self.DT(self.Minv(_lib.PHW_Q, self.D(p)))

    def close(self):
        self.solver.close()
        if self.mesh_handle is not None:
            self.mesh_handle.close()


def calculate_eigen(tFinal, numSteps, outputDir, nx, xLength, yLength, domainShape='R', timeIntScheme='SV', dirichletImp='weak',
                    K_wave=1, rho=1, analytical=False, basis_order=(1, 1), eigen_type='continuous',
                    mesh=None, num_eigs=51, method='auto', device=0):
    """Same call as the reference's calculate_eigen (calculate_eigen.py:14-19).  Returns (out_array [n, 3], numCells) with the columns
    of :222-228: exact eigenvalue, model eigenvalue (positive imaginary part of the pair), real part (zero for this operator)."""
    if eigen_type != 'continuous':
        raise NotImplementedError("only the 'continuous' eigenproblem is restated (the reference marks the discrete one as unchecked, calculate_eigen.py:36-37)")
    if domainShape != 'R' or dirichletImp != 'weak':
        raise NotImplementedError('rectangle domain with weak boundary conditions (calculate_eigen.py:27,31)')
    if mesh is None:
        mesh = generate_mesh(domainShape, nx, xLength, yLength)
    ops = _DeviceOperators(mesh, domainShape, xLength, yLength, basis_order, device)
    c = np.sqrt(K_wave / rho)
    n_p = ops.n_p
    if method == 'auto':
        method = 'dense' if n_p <= 1500 else 'lobpcg'
    if method == 'dense':
        # M_p^-1 D^T M_q^-1 D column by column on the device, then the dense eigenvalues as in the reference (:197-199)
        S = np.empty((n_p, n_p))
        e = np.zeros(n_p)
        for j in range(n_p):
            e[j] = 1.0
            S[:, j] = ops.Minv(_lib.PHW_P, ops.stiffness(e))
            e[j] = 0.0
        lam = np.linalg.eigvals(S)
        lam = lam[np.abs(lam) > 1e-8]
        order = np.argsort(np.abs(lam))
        lam = lam[order]
        omega = c * np.sqrt(np.abs(lam.real))
        real_part = c * np.sqrt(np.abs(lam)) * np.sin(0.5 * np.angle(lam.astype(complex)))      # Re(sqrt(-lam)) of +-i omega: 0 when lam is real positive
    elif method == 'lobpcg':
        import scipy.sparse.linalg as spla
        # two more than asked for: the pencil has the constant pressure as a zero mode, which is dropped below
        k = min(num_eigs + 2, max(1, n_p // 5))
        A = spla.LinearOperator((n_p, n_p), matvec=lambda x: ops.stiffness(np.asarray(x).ravel()), dtype=np.float64)
        B = spla.LinearOperator((n_p, n_p), matvec=lambda x: ops.M(_lib.PHW_P, np.asarray(x).ravel()), dtype=np.float64)
        P = spla.LinearOperator((n_p, n_p), matvec=lambda x: ops.Minv(_lib.PHW_P, np.asarray(x).ravel()), dtype=np.float64)
        rng = np.random.default_rng(0)
        X = rng.standard_normal((n_p, k))
        lam, _ = spla.lobpcg(A, X, B=B, M=P, largest=False, tol=1e-9, maxiter=400)
        lam = np.sort(lam[lam > 1e-8])
        omega = c * np.sqrt(lam)
        real_part = np.zeros_like(omega)
    else:
        raise ValueError("method must be 'auto', 'dense' or 'lobpcg'")
    n = min(num_eigs, omega.shape[0])
    exact = exact_eigenvalues(c, xLength, yLength, n)
    out_array = np.zeros((n, 3))
    out_array[:, 0] = exact
    out_array[:, 1] = omega[:n]
    out_array[:, 2] = real_part[:n]
    numCells = ops.n_cells
    ops.close()
    return out_array, numCells
