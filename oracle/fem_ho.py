"""Oracle finite-element layer for the higher-order pairs P_k x RT_k, k <= 3.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates what dolfin/FFC/FIAT do for ``basis_order = (k_p, k_q)`` of the reference
(``fenics_projects/wave_eqn/wave_2D.py:15-20`` argument, ``:174-175,202-203`` element
definitions ``FiniteElement('P', triangle, k_p)`` / ``FiniteElement('RT', triangle, k_q)``,
forms ``:449-526``; used by the analytical paper cases ``main_wave_2D.py:53-114`` and the
fixtures ``run_test_cases.py:30-31,44-46``).

Independent of the product's host assembly (porthamiltonian_fem_b200/fem_ho.py) on purpose:
  * RT_k is built on the REFERENCE triangle ([P_{k-1}]^2 + x P~_{k-1} in reference
    coordinates) and mapped with the contravariant Piola transform q = J q^ / det J; the
    product builds it directly in scaled physical coordinates of each cell;
  * P_k uses the reference (x, y) node ordering of oracle/fem.py::lagrange_nodes and its own
    cell -> global dof map;
  * quadrature rules of different order.
What both share, by definition (it is the specification of the dofs, needed to compare state
vectors entry by entry; every H_array column is independent of it):
  * P_k dofs = nodal values; numbering [vertices | (k-1) per edge, ordered along the global edge
    direction lo -> hi | (k-1)(k-2)/2 per cell];
  * RT_k dofs = k normal moments per edge  int_e (q.n_e) L_j(s) ds  (L_j Legendre polynomial of
    the global edge parameter s in [0,1], n_e the global edge normal of oracle/fem.py), then
    k(k-1) interior moments per cell  (1/|K|) int_K q_c xi^a eta^b dx  with
    (xi, eta) = (x - x_centroid)/sqrt(2|K|), ordered b-major, a, component fastest;
    numbering [k per edge | k(k-1) per cell].
"""
import numpy as np
import scipy.sparse as sp

from . import fem


def _coo(rows, cols, vals, shape):
    return sp.coo_matrix((np.ravel(vals), (np.ravel(rows), np.ravel(cols))), shape=shape).tocsr()


def _legendre01(j, s):
    return np.polynomial.legendre.Legendre.basis(j)(2.0 * s - 1.0)


def _rt_prime_ref(k, x, y):
    """prime basis of RT_k on the reference triangle at points (x, y):
    values [..., nd, 2] and reference divergences [..., nd]"""
    vals, divs = [], []
    zero = np.zeros_like(x)
    for d in range(k):                       # [P_{k-1}]^2, graded by total degree
        for b in range(d + 1):
            a = d - b
            m = x ** a * y ** b
            mx = (a * x ** (a - 1) if a > 0 else zero) * y ** b
            my = x ** a * (b * y ** (b - 1) if b > 0 else zero)
            vals.append(np.stack([m, zero], -1)); divs.append(mx)
            vals.append(np.stack([zero, m], -1)); divs.append(my)
    for b in range(k):                       # x * homogeneous polynomials of degree k-1
        a = k - 1 - b
        m = x ** a * y ** b
        vals.append(np.stack([x * m, y * m], -1)); divs.append((k + 1) * m)
    return np.stack(vals, -2), np.stack(divs, -1)


class SpacesHO:
    """dof maps + bases of P_kp x RT_kq on a fem.Topology"""

    def __init__(self, topo, kp, kq):
        assert 1 <= kp <= 3 and 1 <= kq <= 3
        self.topo, self.kp, self.kq = topo, kp, kq
        nc, nv, ne = topo.n_cells, topo.n_vertices, topo.n_edges
        V = topo.vertices[topo.cells]                                   # [nc,3,2]
        self.V = V
        self.J = np.stack([V[:, 1] - V[:, 0], V[:, 2] - V[:, 0]], axis=2)   # columns = edge vectors  [nc,2,2]
        self.detJ = self.J[:, 0, 0] * self.J[:, 1, 1] - self.J[:, 0, 1] * self.J[:, 1, 0]
        assert np.all(self.detJ > 0)
        # ---------------- P_k ----------------
        k = kp
        nodes = fem.lagrange_nodes(k)                                   # reference (x, y), y-major
        self.p_nodes_ref = nodes
        ij = np.rint(nodes * k).astype(np.int64)                        # integer lattice (i, j); l0 = k-i-j
        w = np.stack([k - ij[:, 0] - ij[:, 1], ij[:, 0], ij[:, 1]], axis=1)   # integer barycentric weights [nn,3]
        nde, ndi = k - 1, (k - 1) * (k - 2) // 2
        self.ndp = nodes.shape[0]
        self.n_p = nv + ne * nde + nc * ndi
        cd = np.empty((nc, self.ndp), dtype=np.int64)
        n_int = 0
        for n in range(self.ndp):
            nz = np.nonzero(w[n])[0]
            if nz.size == 1:                                            # vertex
                cd[:, n] = topo.cells[:, nz[0]]
            elif nz.size == 2:                                          # on the local edge opposite the zero weight
                e = int(np.nonzero(w[n] == 0)[0][0])
                a, b = nz                                               # local vertices carrying the weights
                ga, gb = topo.cells[:, a], topo.cells[:, b]
                pos = np.where(ga < gb, w[n, b], w[n, a])               # distance from the lower global vertex, in 1/k
                cd[:, n] = nv + topo.cell_edges[:, e] * nde + (pos - 1)
            else:
                cd[:, n] = nv + ne * nde + np.arange(nc) * ndi + n_int
                n_int += 1
        self.cell_dofs_p = cd
        lam = w / float(k)
        xy = np.einsum('nj,cjd->cnd', lam, V)
        self.p_coords = np.zeros((self.n_p, 2))
        self.p_coords[cd.ravel()] = xy.reshape(-1, 2)
        # ---------------- RT_k ----------------
        k = kq
        self.ndq = k * (k + 2)
        nint = k * (k - 1)
        self.n_q = ne * k + nc * nint
        cq = np.empty((nc, self.ndq), dtype=np.int64)
        for e in range(3):
            for j in range(k):
                cq[:, e * k + j] = topo.cell_edges[:, e] * k + j
        for i in range(nint):
            cq[:, 3 * k + i] = ne * k + np.arange(nc) * nint + i
        self.cell_dofs_q = cq
        self.xc = V.mean(axis=1)
        self.h = np.sqrt(2.0 * topo.area)
        # generalised Vandermonde: functional f applied to the Piola-mapped prime function n
        Vm = np.zeros((nc, self.ndq, self.ndq))
        gs, gw = np.polynomial.legendre.leggauss(10)
        s, ws = 0.5 * (gs + 1.0), 0.5 * gw
        for e in range(3):
            a, b = (e + 1) % 3, (e + 2) % 3
            fwd = topo.cell_edge_sign[:, e] > 0                          # cell runs a -> b along lo -> hi
            bary = np.zeros((nc, s.shape[0], 3))
            bary[:, :, a] = np.where(fwd[:, None], 1.0 - s[None, :], s[None, :])
            bary[:, :, b] = 1.0 - bary[:, :, a]
            ge = topo.cell_edges[:, e]
            t = topo.vertices[topo.edges[ge, 1]] - topo.vertices[topo.edges[ge, 0]]
            nL = np.stack([t[:, 1], -t[:, 0]], axis=1)                  # |e| n_e
            val = self.rt_prime_phys(bary)                              # [nc, nq, nd, 2]
            fn = np.einsum('cqnd,cd->cqn', val, nL)
            for j in range(k):
                Vm[:, e * k + j, :] = np.einsum('cqn,q->cn', fn, ws * _legendre01(j, s))
        if nint:
            qx, qy, qw = fem.high_order_tri_rule(10)
            bary = np.broadcast_to(np.stack([1 - qx - qy, qx, qy], axis=1)[None], (nc, qx.shape[0], 3))
            val = self.rt_prime_phys(bary)
            pts = np.einsum('cqj,cjd->cqd', bary, V)
            xi = (pts[..., 0] - self.xc[:, None, 0]) / self.h[:, None]
            eta = (pts[..., 1] - self.xc[:, None, 1]) / self.h[:, None]
            row = 3 * k
            for bb in range(k - 1):
                for aa in range(k - 1 - bb):
                    m = xi ** aa * eta ** bb * qw[None, :]
                    for comp in range(2):
                        Vm[:, row, :] = np.einsum('cqn,cq->cn', val[..., comp], m)
                        row += 1
        self.rt_coef = np.linalg.inv(Vm)                                 # nodal basis_j = sum_n prime_n coef[n, j]

    # RT prime basis mapped to the physical cell (contravariant Piola), at barycentric points [nc, nq, 3]
    def rt_prime_phys(self, bary, with_div=False):
        xh, yh = bary[..., 1], bary[..., 2]
        vhat, dhat = _rt_prime_ref(self.kq, xh, yh)                      # [nc,nq,nd,2], [nc,nq,nd]
        val = np.einsum('cij,cqnj->cqni', self.J, vhat) / self.detJ[:, None, None, None]
        if with_div:
            return val, dhat / self.detJ[:, None, None]
        return val

    def rt_basis(self, bary):
        val, div = self.rt_prime_phys(bary, with_div=True)
        return np.einsum('cqnd,cnj->cqjd', val, self.rt_coef), np.einsum('cqn,cnj->cqj', div, self.rt_coef)

    def p_tabulate(self, x, y):
        return fem.lagrange_tabulate(self.kp, x, y)                     # [npts, ndp], reference points

    # ------------------------------------------------------------------------------------------
    def assemble(self, tags, neumann_tags=(fem.TAG_TOP, fem.TAG_GENERAL), left_weight=None):
        """M_p, M_q, C, N, D = C - N, b_L of wave_2D.py:449-518 for this pair."""
        topo = self.topo
        nc = topo.n_cells
        qx, qy, qw = fem.high_order_tri_rule(9)
        psi = self.p_tabulate(qx, qy)                                    # [nq, ndp]
        bary = np.broadcast_to(np.stack([1 - qx - qy, qx, qy], axis=1)[None], (nc, qx.shape[0], 3))
        phi, div = self.rt_basis(bary)
        w = qw[None, :] * topo.area[:, None]
        Mp = np.einsum('qi,qj,q->ij', psi, psi, qw)[None] * topo.area[:, None, None]
        Mq = np.einsum('cqid,cqjd,cq->cij', phi, phi, w)
        C = np.einsum('cqi,qj,cq->cij', div, psi, w)
        cp, cq = self.cell_dofs_p, self.cell_dofs_q
        ops = {
            'M_p': _coo(np.repeat(cp[:, :, None], self.ndp, 2), np.repeat(cp[:, None, :], self.ndp, 1), Mp, (self.n_p, self.n_p)),
            'M_q': _coo(np.repeat(cq[:, :, None], self.ndq, 2), np.repeat(cq[:, None, :], self.ndq, 1), Mq, (self.n_q, self.n_q)),
            'C': _coo(np.repeat(cq[:, :, None], self.ndp, 2), np.repeat(cp[:, None, :], self.ndq, 1), C, (self.n_q, self.n_p)),
        }
        # exterior facets
        etag = np.full(topo.n_edges, -1, dtype=np.int64)
        etag[topo.boundary_edges] = tags
        gs, gw = np.polynomial.legendre.leggauss(10)
        s, ws = 0.5 * (gs + 1.0), 0.5 * gw
        rows, cols, vals = [], [], []
        bL = np.zeros(self.n_q)
        for e in range(3):
            cidx = np.nonzero(etag[topo.cell_edges[:, e]] >= 0)[0]
            if cidx.size == 0:
                continue
            a, b = (e + 1) % 3, (e + 2) % 3
            pa, pb = self.V[cidx, a], self.V[cidx, b]
            t = pb - pa
            L = np.hypot(t[:, 0], t[:, 1])
            nout = np.stack([t[:, 1], -t[:, 0]], axis=1) / L[:, None]
            bary_e = np.zeros((nc, s.shape[0], 3))
            bary_e[:, :, a] = 1.0 - s[None, :]
            bary_e[:, :, b] = s[None, :]
            phi_e, _ = self.rt_basis(bary_e)
            phin = np.einsum('cqjd,cd->cqj', phi_e[cidx], nout)                  # [nb, nq, ndq]
            psi_e = self.p_tabulate(bary_e[0, :, 1], bary_e[0, :, 2])            # [nq, ndp]
            tg = etag[topo.cell_edges[cidx, e]]
            neu = np.isin(tg, neumann_tags)
            if neu.any():
                Nloc = np.einsum('cqi,qj,q->cij', phin[neu], psi_e, ws) * L[neu, None, None]
                rows.append(np.repeat(cq[cidx[neu]][:, :, None], self.ndp, 2).ravel())
                cols.append(np.repeat(cp[cidx[neu]][:, None, :], self.ndq, 1).ravel())
                vals.append(Nloc.ravel())
            left = tg == fem.TAG_LEFT
            if left.any():
                if left_weight is None:
                    gq = np.ones((int(left.sum()), s.shape[0]))
                else:
                    # degree-8 Expression (wave_2D.py:376-377): its trace on a facet is the degree-8 interpolant at
                    # 9 equispaced points; evaluated at the quadrature points with the barycentric formula
                    t9 = np.linspace(0.0, 1.0, 9)
                    p9 = pa[left][:, None, :] + t9[None, :, None] * t[left][:, None, :]
                    g9 = left_weight(p9[..., 0], p9[..., 1])
                    wb = np.array([(-1.0) ** j * _binom(8, j) for j in range(9)])
                    num = (g9 * wb)[:, None, :] / (s[None, :, None] - t9[None, None, :])
                    den = wb[None, None, :] / (s[None, :, None] - t9[None, None, :])
                    gq = num.sum(-1) / den.sum(-1)
                contrib = np.einsum('cqi,cq,q->ci', phin[left], gq, ws) * L[left, None]
                np.add.at(bL, cq[cidx[left]].ravel(), contrib.ravel())
        if rows:
            N = _coo(np.concatenate(rows), np.concatenate(cols), np.concatenate(vals), (self.n_q, self.n_p))
        else:
            N = sp.csr_matrix((self.n_q, self.n_p))
        ops['N'] = N
        ops['D'] = (ops['C'] - N).tocsr()
        ops['b_L'] = bL
        return ops


def _binom(n, j):
    from math import comb
    return float(comb(n, j))


class AnalyticalDiagnosticsHO:
    """wave_2D.py:929-965 for a P_k field: errornorm with degree rise 3 (:956), max nodal error over the P_k dofs
    (:958), point probe (:946-950)."""

    def __init__(self, spaces, rho_val, omega, xLength, yLength, xPoint=0.155, yPoint=0.189):
        topo = spaces.topo
        self.sp = spaces
        self.omega = omega
        self.space = lambda x, y: rho_val * omega * np.sin(np.pi * x / (2 * xLength) + np.pi / 2) * \
            np.sin(2 * np.pi * y / yLength + np.pi / 2)
        kk = spaces.kp + 3
        nodes = fem.lagrange_nodes(kk)
        lam = np.stack([1 - nodes[:, 0] - nodes[:, 1], nodes[:, 0], nodes[:, 1]], axis=1)
        xy = np.einsum('nj,cjd->cnd', lam, spaces.V)
        self.exact_nodes = self.space(xy[..., 0], xy[..., 1])
        self.T = spaces.p_tabulate(nodes[:, 0], nodes[:, 1])                      # P_k basis at the P_{k+3} nodes
        qx, qy, qw = fem.high_order_tri_rule(8)
        tab = fem.lagrange_tabulate(kk, qx, qy)
        self.mass_ref = np.einsum('qi,qj,q->ij', tab, tab, qw)
        self.exact_dofs = self.space(spaces.p_coords[:, 0], spaces.p_coords[:, 1])
        P = np.array([xPoint, yPoint])
        rhs = P[None, :] - spaces.V[:, 0]
        Jm = spaces.J
        l1 = (rhs[:, 0] * Jm[:, 1, 1] - rhs[:, 1] * Jm[:, 0, 1]) / spaces.detJ
        l2 = (Jm[:, 0, 0] * rhs[:, 1] - Jm[:, 1, 0] * rhs[:, 0]) / spaces.detJ
        l0 = 1 - l1 - l2
        inside = np.nonzero((l0 >= -1e-12) & (l1 >= -1e-12) & (l2 >= -1e-12))[0]
        self.pt_cell = int(inside[0]) if inside.size else -1
        if self.pt_cell >= 0:
            self.pt_w = spaces.p_tabulate(np.array([l1[self.pt_cell]]), np.array([l2[self.pt_cell]]))[0]

    def evaluate(self, p_h, t):
        ct = np.cos(t * self.omega)
        ph_nodes = p_h[self.sp.cell_dofs_p] @ self.T.T
        e = self.exact_nodes * ct - ph_nodes
        err_L2 = np.sqrt(np.sum(np.einsum('ci,ij,cj->c', e, self.mass_ref, e) * self.sp.topo.area))
        err_max = np.abs(self.exact_dofs * ct - p_h).max()
        if self.pt_cell >= 0:
            d = self.sp.cell_dofs_p[self.pt_cell]
            return err_L2, err_max, float(p_h[d] @ self.pt_w), float((self.exact_dofs[d] * ct) @ self.pt_w)
        return err_L2, err_max, 0.0, 0.0
