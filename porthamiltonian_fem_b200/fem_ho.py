"""Higher-order element pairs P_k x RT_k, k <= 3 (reference: ``basis_order`` of wave_2D.py:15-20,174-175,202-203;
used by the 'analytical' paper cases, main_wave_2D.py:53-114).

Host-side, one-time setup (numpy): dof maps, element tabulations and the assembled sparse operators of the forms
of wave_2D.py:449-526 for arbitrary order; the per-step work then runs on the device through the CSR kernels of
libphwave (phw_ho_*).  For the lowest-order pair the matrix-free kernels are used instead.

Construction (all outputs of the solver are basis independent, so any conforming basis of the same spaces gives the
reference's H_array):
  * P_k: Lagrange basis on the equispaced lattice; global dofs = vertices, (k-1) per edge ordered along the global
    edge direction lo -> hi, (k-1)(k-2)/2 per cell.
  * RT_k (FEniCS 'RT' degree k: [P_{k-1}]^2 + x P~_{k-1}, k(k+2) dofs per cell): dual basis of the functionals
      - k normal moments per edge:  int_e (q.n_e) L_j(s) ds, L_j = Legendre polynomial in the global edge parameter,
        n_e the global edge normal (so neighbouring cells share the functional => normal continuity),
      - k(k-1) interior moments against a monomial basis of [P_{k-2}]^2,
    computed cell by cell in the physical cell (generalised Vandermonde inversion, batched).
"""
import numpy as np
import scipy.sparse as sp

from .partition import edge_table


# --------------------------------------------------------------------------------------------------------------
# quadrature
# --------------------------------------------------------------------------------------------------------------
def tri_rule(n=7):
    """collapsed Gauss-Legendre rule, n*n points, exact to degree 2n-2; barycentric points [nq,3], weights sum to 1"""
    gx, gw = np.polynomial.legendre.leggauss(n)
    u, wu = 0.5 * (gx + 1), 0.5 * gw
    U, V = np.meshgrid(u, u, indexing='ij')
    x = U.ravel()
    y = (V * (1 - U)).ravel()
    w = 2.0 * (np.outer(wu, wu) * (1 - U)).ravel()
    return np.stack([1 - x - y, x, y], axis=1), w


def edge_rule(n=8):
    gx, gw = np.polynomial.legendre.leggauss(n)
    return 0.5 * (gx + 1), 0.5 * gw


def legendre01(j, s):
    """Legendre polynomial P_j on [0,1]"""
    return np.polynomial.legendre.Legendre.basis(j)(2 * s - 1)


# --------------------------------------------------------------------------------------------------------------
# P_k Lagrange
# --------------------------------------------------------------------------------------------------------------
def lattice(k):
    """equispaced lattice in barycentric integer coordinates (i0,i1,i2), i0+i1+i2 = k, grouped as
    [3 vertices | edge 0 | edge 1 | edge 2 | interior]; edge i runs from local vertex i+1 to i+2."""
    pts = []
    for v in range(3):
        p = [0, 0, 0]
        p[v] = k
        pts.append(tuple(p))
    for e in range(3):
        a, b = (e + 1) % 3, (e + 2) % 3
        for j in range(1, k):
            p = [0, 0, 0]
            p[a], p[b] = k - j, j
            pts.append(tuple(p))
    for i1 in range(1, k):
        for i2 in range(1, k - i1):
            pts.append((k - i1 - i2, i1, i2))
    return np.array(pts, dtype=np.int64)


def _mono(k, l1, l2):
    return np.stack([l1 ** a * l2 ** b for b in range(k + 1) for a in range(k + 1 - b)], axis=-1)


def _dmono(k, l1, l2):
    d1 = np.stack([(a * l1 ** (a - 1) if a > 0 else 0 * l1) * l2 ** b for b in range(k + 1) for a in range(k + 1 - b)], axis=-1)
    d2 = np.stack([l1 ** a * (b * l2 ** (b - 1) if b > 0 else 0 * l2) for b in range(k + 1) for a in range(k + 1 - b)], axis=-1)
    return d1, d2


class LagrangeRef:
    """P_k nodal basis on the reference triangle in the lattice ordering above."""

    def __init__(self, k):
        self.k = k
        self.lat = lattice(k)
        lam = self.lat / float(k)
        self.nodes_bary = lam
        self.coef = np.linalg.inv(_mono(k, lam[:, 1], lam[:, 2]))      # monomial coefficients of each basis function

    def tabulate(self, bary):
        return _mono(self.k, bary[:, 1], bary[:, 2]) @ self.coef        # [npts, ndof]


# --------------------------------------------------------------------------------------------------------------
# the discretisation object
# --------------------------------------------------------------------------------------------------------------
class HighOrderSpaces:
    def __init__(self, vertices, cells, kp, kq):
        if not (1 <= kp <= 3 and 1 <= kq <= 3):
            raise NotImplementedError('basis orders 1..3 are supported')
        self.kp, self.kq = kp, kq
        vertices = np.ascontiguousarray(vertices, dtype=np.float64)
        cells = np.array(cells, dtype=np.int64)
        V = vertices[cells]
        area2 = (V[:, 1, 0] - V[:, 0, 0]) * (V[:, 2, 1] - V[:, 0, 1]) - (V[:, 1, 1] - V[:, 0, 1]) * (V[:, 2, 0] - V[:, 0, 0])
        flip = area2 < 0
        cells[flip, 1], cells[flip, 2] = cells[flip, 2].copy(), cells[flip, 1].copy()
        self.vertices, self.cells = vertices, cells
        self.V = vertices[cells]
        self.area = 0.5 * np.abs(area2)
        self.nv, self.nc = vertices.shape[0], cells.shape[0]
        self.edges, self.cell_edges = edge_table(cells, self.nv)
        self.ne = self.edges.shape[0]
        a = cells[:, [1, 2, 0]]
        b = cells[:, [2, 0, 1]]
        self.cell_edge_sign = np.where(a < b, 1, -1)
        cnt = np.bincount(self.cell_edges.ravel(), minlength=self.ne)
        self.boundary_edges = np.nonzero(cnt == 1)[0]
        self._build_p()
        self._build_q()

    # ---- P_k ---------------------------------------------------------------------------------------------
    def _build_p(self):
        k = self.kp
        self.lag = LagrangeRef(k)
        nde, ndi = k - 1, (k - 1) * (k - 2) // 2
        self.ndp = 3 + 3 * nde + ndi
        self.np_dofs = self.nv + self.ne * nde + self.nc * ndi
        cd = np.empty((self.nc, self.ndp), dtype=np.int64)
        cd[:, :3] = self.cells
        for e in range(3):
            ge = self.cell_edges[:, e]
            fwd = self.cell_edge_sign[:, e] > 0
            for j in range(1, k):          # local point j along local direction; global index along lo -> hi
                gj = np.where(fwd, j, k - j)
                cd[:, 3 + e * nde + (j - 1)] = self.nv + ge * nde + (gj - 1)
        for i in range(ndi):
            cd[:, 3 + 3 * nde + i] = self.nv + self.ne * nde + np.arange(self.nc) * ndi + i
        self.cell_dofs_p = cd
        # dof coordinates (nodal interpolation points)
        xy = np.einsum('nj,cjd->cnd', self.lag.nodes_bary, self.V)
        coords = np.zeros((self.np_dofs, 2))
        coords[cd.ravel()] = xy.reshape(-1, 2)
        self.p_coords = coords

    # ---- RT_k --------------------------------------------------------------------------------------------
    def _prime(self, xi, eta):
        """prime basis of RT_k in scaled cell coordinates: returns [..., nd, 2] values and [..., nd] divergences (d/dxi)"""
        k = self.kq
        fx, fy, dv = [], [], []
        for b in range(k):
            for a in range(k - b):           # monomials of degree <= k-1
                m = xi ** a * eta ** b
                dmx = (a * xi ** (a - 1) if a > 0 else 0 * xi) * eta ** b
                dmy = xi ** a * (b * eta ** (b - 1) if b > 0 else 0 * eta)
                fx.append(m); fy.append(0 * m); dv.append(dmx)
                fx.append(0 * m); fy.append(m); dv.append(dmy)
        for a in range(k):                   # x * homogeneous degree k-1
            b = k - 1 - a
            m = xi ** a * eta ** b
            fx.append(xi * m); fy.append(eta * m); dv.append((a + b + 2) * m)
        return np.stack([np.stack(fx, -1), np.stack(fy, -1)], -1), np.stack(dv, -1)

    def _build_q(self):
        k = self.kq
        self.ndq = k * (k + 2)
        n_int = k * (k - 1)
        self.nq_dofs = self.ne * k + self.nc * n_int
        cd = np.empty((self.nc, self.ndq), dtype=np.int64)
        for e in range(3):
            for j in range(k):
                cd[:, e * k + j] = self.cell_edges[:, e] * k + j
        for i in range(n_int):
            cd[:, 3 * k + i] = self.ne * k + np.arange(self.nc) * n_int + i
        self.cell_dofs_q = cd
        # scaled coordinates: xi = (x - xc) / h
        self.xc = self.V.mean(axis=1)
        self.h = np.sqrt(2.0 * self.area)
        # functionals applied to the prime basis -> Vandermonde [nc, nfun, nd]
        Vm = np.zeros((self.nc, self.ndq, self.ndq))
        s, w = edge_rule(8)
        for e in range(3):
            ge = self.cell_edges[:, e]
            plo, phi = self.vertices[self.edges[ge, 0]], self.vertices[self.edges[ge, 1]]
            t = phi - plo
            nrm = np.stack([t[:, 1], -t[:, 0]], axis=1)                       # |e| n_e (global normal)
            pts = plo[:, None, :] + s[None, :, None] * t[:, None, :]          # [nc, nq, 2] along lo -> hi
            xi = (pts[..., 0] - self.xc[:, None, 0]) / self.h[:, None]
            eta = (pts[..., 1] - self.xc[:, None, 1]) / self.h[:, None]
            f, _ = self._prime(xi, eta)                                       # [nc, nq, nd, 2]
            fn = np.einsum('cqnd,cd->cqn', f, nrm)                            # (q.n_e)|e| at the points
            for j in range(k):
                Vm[:, e * k + j, :] = np.einsum('cqn,q->cn', fn, w * legendre01(j, s))
        if n_int:
            bary, wq = tri_rule(7)
            pts = np.einsum('qj,cjd->cqd', bary, self.V)
            xi = (pts[..., 0] - self.xc[:, None, 0]) / self.h[:, None]
            eta = (pts[..., 1] - self.xc[:, None, 1]) / self.h[:, None]
            f, _ = self._prime(xi, eta)
            row = 3 * k
            for b in range(k - 1):
                for a in range(k - 1 - b):
                    m = xi ** a * eta ** b
                    for comp in range(2):
                        Vm[:, row, :] = np.einsum('cqn,cq->cn', f[..., comp], m * wq[None, :])
                        row += 1
            assert row == self.ndq
        self.rt_coef = np.linalg.inv(Vm)          # basis_j = sum_n prime_n coef[n, j]

    def rt_tabulate(self, pts):
        """RT basis at physical points pts [nc, nq, 2]: values [nc, nq, nd, 2], divergence [nc, nq, nd]"""
        xi = (pts[..., 0] - self.xc[:, None, 0]) / self.h[:, None]
        eta = (pts[..., 1] - self.xc[:, None, 1]) / self.h[:, None]
        f, dv = self._prime(xi, eta)
        val = np.einsum('cqnd,cnj->cqjd', f, self.rt_coef)
        div = np.einsum('cqn,cnj->cqj', dv, self.rt_coef) / self.h[:, None, None]
        return val, div

    # ---- assembly ----------------------------------------------------------------------------------------
    def assemble(self, btags_by_edge, left_weight=None, neumann_tags=(1, 3)):
        """btags_by_edge: dict edge id -> tag for boundary edges.  Returns dict of CSR operators
        M_p, M_q, D = C - N, b_L (vector, weight g_L = left_weight(x,y) or 1), plus b_L1 (weight 1)."""
        bary, wq = tri_rule(7)
        psi = self.lag.tabulate(bary)                                          # [nq, ndp]
        pts = np.einsum('qj,cjd->cqd', bary, self.V)
        phi, div = self.rt_tabulate(pts)
        w = wq[None, :] * self.area[:, None]
        Mp = np.einsum('qi,qj,cq->cij', psi, psi, w)
        Mq = np.einsum('cqid,cqjd,cq->cij', phi, phi, w)
        C = np.einsum('cqi,qj,cq->cij', div, psi, w)

        def coo(rows, cols, vals, shape):
            return sp.coo_matrix((vals.ravel(), (rows.ravel(), cols.ravel())), shape=shape).tocsr()
        rp = np.repeat(self.cell_dofs_p[:, :, None], self.ndp, axis=2)
        cp = np.repeat(self.cell_dofs_p[:, None, :], self.ndp, axis=1)
        rq = np.repeat(self.cell_dofs_q[:, :, None], self.ndq, axis=2)
        cq = np.repeat(self.cell_dofs_q[:, None, :], self.ndq, axis=1)
        rqp = np.repeat(self.cell_dofs_q[:, :, None], self.ndp, axis=2)
        cqp = np.repeat(self.cell_dofs_p[:, None, :], self.ndq, axis=1)
        ops = {'M_p': coo(rp, cp, Mp, (self.np_dofs, self.np_dofs)),
               'M_q': coo(rq, cq, Mq, (self.nq_dofs, self.nq_dofs)),
               'C': coo(rqp, cqp, C, (self.nq_dofs, self.np_dofs))}
        # boundary facets
        s, ws = edge_rule(8)
        etag = -np.ones(self.ne, dtype=np.int64)
        for e, t in btags_by_edge.items():
            etag[e] = t
        Nrows, Ncols, Nvals = [], [], []
        bL = np.zeros(self.nq_dofs)
        bL1 = np.zeros(self.nq_dofs)
        for e in range(3):
            ge = self.cell_edges[:, e]
            isb = etag[ge] >= 0
            cidx = np.nonzero(isb)[0]
            if cidx.size == 0:
                continue
            a = self.V[cidx, (e + 1) % 3]
            b = self.V[cidx, (e + 2) % 3]
            t = b - a
            L = np.hypot(t[:, 0], t[:, 1])
            nout = np.stack([t[:, 1], -t[:, 0]], axis=1) / L[:, None]          # outward normal of a CCW cell
            pts_e = a[:, None, :] + s[None, :, None] * t[:, None, :]
            xi = (pts_e[..., 0] - self.xc[cidx, None, 0]) / self.h[cidx, None]
            eta = (pts_e[..., 1] - self.xc[cidx, None, 1]) / self.h[cidx, None]
            f, _ = self._prime(xi, eta)
            val = np.einsum('cqnd,cnj->cqjd', f, self.rt_coef[cidx])
            phin = np.einsum('cqjd,cd->cqj', val, nout)                        # [nb, nq, ndq]
            bar = np.zeros((s.shape[0], 3))
            bar[:, (e + 1) % 3] = 1 - s
            bar[:, (e + 2) % 3] = s
            psi_e = self.lag.tabulate(bar)                                     # [nq, ndp]
            tags = etag[ge[cidx]]
            neu = np.isin(tags, neumann_tags)
            if neu.any():
                Nloc = np.einsum('cqi,qj,q,c->cij', phin[neu], psi_e, ws, L[neu])
                Nrows.append(np.repeat(self.cell_dofs_q[cidx[neu]][:, :, None], self.ndp, axis=2).ravel())
                Ncols.append(np.repeat(self.cell_dofs_p[cidx[neu]][:, None, :], self.ndq, axis=1).ravel())
                Nvals.append(Nloc.ravel())
            left = tags == 0
            if left.any():
                one = np.einsum('cqi,q,c->ci', phin[left], ws, L[left])
                np.add.at(bL1, self.cell_dofs_q[cidx[left]].ravel(), one.ravel())
                if left_weight is None:
                    np.add.at(bL, self.cell_dofs_q[cidx[left]].ravel(), one.ravel())
                else:
                    # degree-8 interpolant of the datum on the facet (Expression degree=8, wave_2D.py:376-377)
                    t9 = np.linspace(0.0, 1.0, 9)
                    p9 = a[left][:, None, :] + t9[None, :, None] * t[left][:, None, :]
                    g9 = left_weight(p9[..., 0], p9[..., 1])                  # [nl, 9]
                    Lmat = _lagrange_matrix(t9, s)                            # [nq, 9]
                    gq = g9 @ Lmat.T
                    wgt = np.einsum('cqi,cq,q,c->ci', phin[left], gq, ws, L[left])
                    np.add.at(bL, self.cell_dofs_q[cidx[left]].ravel(), wgt.ravel())
        if Nrows:
            N = sp.coo_matrix((np.concatenate(Nvals), (np.concatenate(Nrows), np.concatenate(Ncols))),
                              shape=(self.nq_dofs, self.np_dofs)).tocsr()
        else:
            N = sp.csr_matrix((self.nq_dofs, self.np_dofs))
        ops['N'] = N
        D = (ops['C'] - N).tocsr()
        D.eliminate_zeros()
        ops['D'] = D
        ops['b_L'] = bL
        ops['b_L1'] = bL1
        return ops

    # ---- spectral bounds of the Jacobi-scaled mass matrices (element-wise, Wathen) -------------------------
    def jacobi_bounds(self, ops):
        def bounds(M, cell_dofs, nd, Mloc):
            d = M.diagonal()
            lo, hi = np.inf, 0.0
            # generalised eigenvalues of (M_K, diag-part D_K of the *assembled* diagonal restricted to K is not
            # element-local; use the element's own diagonal (Wathen's bound))
            dk = np.einsum('cii->ci', Mloc)
            S = Mloc / np.sqrt(dk[:, :, None] * dk[:, None, :])
            ev = np.linalg.eigvalsh(S)
            return float(ev.min()), float(ev.max())
        bary, wq = tri_rule(7)
        psi = self.lag.tabulate(bary)
        w = wq[None, :] * self.area[:, None]
        Mp = np.einsum('qi,qj,cq->cij', psi, psi, w[:1])       # affine invariant up to |K|: one cell suffices
        pts = np.einsum('qj,cjd->cqd', bary, self.V)
        phi, _ = self.rt_tabulate(pts)
        Mq = np.einsum('cqid,cqjd,cq->cij', phi, phi, w)
        return bounds(ops['M_p'], self.cell_dofs_p, self.ndp, Mp), bounds(ops['M_q'], self.cell_dofs_q, self.ndq, Mq)

    # ---- helpers for initial data and diagnostics -----------------------------------------------------------
    def interpolate_p(self, fun_xy):
        return fun_xy(self.p_coords[:, 0], self.p_coords[:, 1])


def _lagrange_matrix(nodes, x):
    """values at x of the Lagrange basis polynomials on `nodes`: [len(x), len(nodes)]"""
    out = np.ones((x.shape[0], nodes.shape[0]))
    for j in range(nodes.shape[0]):
        for m in range(nodes.shape[0]):
            if m != j:
                out[:, j] *= (x - nodes[m]) / (nodes[j] - nodes[m])
    return out
