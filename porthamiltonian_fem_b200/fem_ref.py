"""Reference-element form of the higher-order pairs P_k x RT_k, k <= 3 (``basis_order`` of wave_2D.py:15-20,174-175,202-203): what
the matrix-free device kernels of csrc/kernels_ho.cuh work with.

Every cell is taken with its vertices in ascending global order and mapped from the reference triangle (0,0), (1,0), (0,1) by
x = x0 + J xhat.  With that ordering the three local edges (v1->v2, v0->v2, v0->v1) run along the global edge direction lo -> hi in
every cell, so no per-cell orientation flips are needed, and with the bases

    P_k : Lagrange on the equispaced lattice,                          psi  = psihat o F^-1
    RT_k: dual to  k normal moments per edge against Legendre polynomials of the edge parameter,  int_e (q . rot t) L_j ds,
          and k(k-1) interior (bubble) functions orthonormal in the mass product of an undistorted cell,  phi = J phihat / det J  (Piola),

the element matrices of the forms wave_2D.py:449-526 are REFERENCE matrices times a few geometric numbers per cell:

    M_p,K = |det J| Mp_hat        C_K = sign(det J) C_hat        M_q,K = G00 A00 + G01 (A01 + A10) + G11 A11,   G = J^T J / |det J|

and the boundary terms act on the k dofs of a boundary edge only (their normal trace is (2j+1) L_j(s) / |e|, every other basis
function has none): N_e = sigma_e Nhat, b_L,e = sigma_e int (2j+1) L_j g_L, B_top,e = |e| (1-D mass), B_right,e = diag(2j+1) / |e|.
The device keeps 3 vertex ids + 3 edge ids + 4 doubles per cell and the reference tabulations in shared memory; nothing is assembled.

Dof numbering and meaning are those of fem_ho.HighOrderSpaces / oracle.fem_ho (the specification both share): P_k nodal values
[vertices | k-1 per edge along lo -> hi | interior], RT_k [k moments per edge | k(k-1) interior moments per cell].  The edge moments
are the same functionals here; the INTERIOR moments of the specification are taken against monomials of scaled physical
coordinates, not reference coordinates, so the interior dofs of the two bases differ by a cell-local linear map: ``to_spec`` /
``from_spec`` convert state vectors at the API boundary (set_state / get_state), the device works in the reference basis.
"""
import numpy as np
import scipy.sparse as sp

from .fem_ho import tri_rule, edge_rule, legendre01, _lagrange_matrix
from .partition import edge_table

LOCAL_EDGES = ((1, 2), (0, 2), (0, 1))      # local edge e is opposite local vertex e and runs from the lower to the higher vertex


def _mono2(k, x, y):
    """monomials x^a y^b, a + b <= k (b-major)"""
    return np.stack([x ** a * y ** b for b in range(k + 1) for a in range(k + 1 - b)], axis=-1)


def _rt_prime(k, x, y):
    """prime basis of RT_k = [P_{k-1}]^2 + x P~_{k-1}: values [..., nd, 2] and divergences [..., nd]"""
    fx, fy, dv = [], [], []
    for b in range(k):
        for a in range(k - b):
            m = x ** a * y ** b
            dmx = (a * x ** (a - 1) if a > 0 else 0 * x) * y ** b
            dmy = x ** a * (b * y ** (b - 1) if b > 0 else 0 * y)
            fx.append(m); fy.append(0 * m); dv.append(dmx)
            fx.append(0 * m); fy.append(m); dv.append(dmy)
    for a in range(k):
        b = k - 1 - a
        m = x ** a * y ** b
        fx.append(x * m); fy.append(y * m); dv.append((a + b + 2) * m)
    return np.stack([np.stack(fx, -1), np.stack(fy, -1)], -1), np.stack(dv, -1)


class RefElement:
    """reference tabulations of the pair (kp, kq)"""

    def __init__(self, kp, kq):
        if not (1 <= kp <= 3 and 1 <= kq <= 3):
            raise NotImplementedError('basis orders 1..3 are supported')
        self.kp, self.kq = kp, kq
        vref = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]])
        # ---- P_k lattice in the local ordering [vertices | edge 0 | edge 1 | edge 2 | interior]
        k = kp
        pts = [tuple(v) for v in vref]
        for (a, b) in LOCAL_EDGES:
            for j in range(1, k):
                pts.append(tuple(vref[a] + (j / k) * (vref[b] - vref[a])))
        for i1 in range(1, k):
            for i2 in range(1, k - i1):
                pts.append((i1 / k, i2 / k))
        self.p_nodes = np.array(pts)                                            # reference (x, y)
        self.ndp = self.p_nodes.shape[0]
        self.p_coef = np.linalg.inv(_mono2(k, self.p_nodes[:, 0], self.p_nodes[:, 1]))
        # ---- RT_k nodal basis on the reference triangle
        k = kq
        self.ndq, self.n_int = k * (k + 2), k * (k - 1)
        Vm = np.zeros((self.ndq, self.ndq))
        s, ws = edge_rule(10)
        for e, (a, b) in enumerate(LOCAL_EDGES):
            t = vref[b] - vref[a]
            rot = np.array([t[1], -t[0]])
            p = vref[a][None, :] + s[:, None] * t[None, :]
            f, _ = _rt_prime(k, p[:, 0], p[:, 1])                               # [nq, nd, 2]
            fn = f @ rot
            for j in range(k):
                Vm[e * k + j] = (ws * legendre01(j, s)) @ fn
        bary, wq = tri_rule(8)
        self.q_bary, self.q_w = bary, 0.5 * wq                                  # rule on the reference triangle (area 1/2)
        xh, yh = bary[:, 1], bary[:, 2]
        f, dv = _rt_prime(k, xh, yh)
        row = 3 * k
        self.int_coef = self._interior_polys(k, xh, yh)                         # test polynomials of the interior moments
        mono = _mono2(max(k - 2, 0), xh, yh)
        for i in range(self.int_coef.shape[1] if k > 1 else 0):
            m = (mono @ self.int_coef[:, i]) * self.q_w
            for comp in range(2):
                Vm[row] = m @ f[..., comp]
                row += 1
        self.q_coef = np.linalg.inv(Vm)                                         # basis_j = sum_n prime_n coef[n, j]
        if self.n_int:
            # The interior (bubble) functions are cell-local, so any basis of their span will do.  Taken orthonormal in the mass
            # inner product of an undistorted cell (G = I: A00 + A11), with the edge functions made orthogonal to them, the element
            # mass matrix of a right-isosceles cell is block diagonal with an identity interior block: the Jacobi-scaled spectrum
            # (Chebyshev interval of the M_q solve) is then set by the edge block alone.  Edge moments are untouched (bubbles have none).
            phi0, _ = self.tab_q(xh, yh)
            M0 = np.einsum('qid,qjd,q->ij', phi0, phi0, self.q_w)
            ke = 3 * k
            Mii, Mie = M0[ke:, ke:], M0[ke:, :ke]
            Tc = np.eye(self.ndq)
            Tc[ke:, :ke] = -np.linalg.solve(Mii, Mie)
            Tc[ke:, ke:] = np.linalg.inv(np.linalg.cholesky(Mii)).T
            self.q_coef = self.q_coef @ Tc
        # ---- reference matrices
        psi = self.tab_p(xh, yh)                                                # [nq, ndp]
        phi, div = self.tab_q(xh, yh)                                           # [nq, ndq, 2], [nq, ndq]
        w = self.q_w
        self.Mp_hat = np.einsum('qi,qj,q->ij', psi, psi, w)
        self.C_hat = np.einsum('qi,qj,q->ij', div, psi, w)                      # [ndq, ndp]
        self.A00 = np.einsum('qi,qj,q->ij', phi[..., 0], phi[..., 0], w)
        self.A11 = np.einsum('qi,qj,q->ij', phi[..., 1], phi[..., 1], w)
        A01 = np.einsum('qi,qj,q->ij', phi[..., 0], phi[..., 1], w)
        self.A01s = A01 + A01.T
        # ---- one boundary edge: P_k trace = 1-D Lagrange basis on [lo, hi, k-1 interior points], RT_k normal trace (2j+1) L_j(s)
        s1, w1 = edge_rule(10)
        nodes1 = np.concatenate([[0.0, 1.0], np.arange(1, kp) / kp])
        self.edge_nodes = nodes1
        L1 = _lagrange_matrix(nodes1, s1)                                       # [nq, kp+1]
        tr = np.stack([(2 * j + 1) * legendre01(j, s1) for j in range(kq)], axis=1)   # [nq, kq]
        self.N_hat = np.einsum('qj,qm,q->jm', tr, L1, w1)                       # <phi_{e,j}.n |e|, psi_m> / sigma
        self.B1_hat = np.einsum('qm,qn,q->mn', L1, L1, w1)                      # 1-D P_k mass on the unit interval
        self.trace_q, self.edge_s, self.edge_w = tr, s1, w1

    def _interior_polys(self, k, xh, yh):
        """L2(Khat)-orthonormal basis of P_{k-2} (Gram-Schmidt of the monomials): the interior moments are taken against these,
        component by component - far better conditioned after Jacobi scaling than raw monomials (spectral bounds of M_q)"""
        if k < 2:
            return np.zeros((1, 0))
        mono = _mono2(k - 2, xh, yh)
        G = np.einsum('qi,qj,q->ij', mono, mono, self.q_w)
        return np.linalg.inv(np.linalg.cholesky(G)).T                          # columns = coefficients of the orthonormal polynomials

    def interior_test(self, xh, yh):
        """values of the interior test polynomials at reference points [..., n_int / 2]"""
        return _mono2(max(self.kq - 2, 0), xh, yh) @ self.int_coef

    def tab_p(self, x, y):
        return _mono2(self.kp, x, y) @ self.p_coef

    def tab_q(self, x, y):
        f, dv = _rt_prime(self.kq, x, y)
        return np.einsum('...nd,nj->...jd', f, self.q_coef), dv @ self.q_coef


class RefSpaces:
    """dof maps, per-cell geometry and boundary operators of P_kp x RT_kq on a triangle mesh, in the reference-element form"""

    def __init__(self, vertices, cells, kp, kq):
        self.ref = RefElement(kp, kq)
        self.kp, self.kq = kp, kq
        self.vertices = np.ascontiguousarray(vertices, dtype=np.float64)
        self.cells_in = np.asarray(cells, dtype=np.int64)
        self.cells = np.sort(self.cells_in, axis=1)                             # ascending global vertex ids
        self.nv, self.nc = self.vertices.shape[0], self.cells.shape[0]
        self.edges, ce = edge_table(self.cells, self.nv)                        # ce[:, i] = edge between local vertices i+1, i+2
        self.cell_edges = ce                                                    # = edge opposite local vertex i  (as LOCAL_EDGES)
        self.ne = self.edges.shape[0]
        V = self.vertices[self.cells]
        self.V = V
        J = np.stack([V[:, 1] - V[:, 0], V[:, 2] - V[:, 0]], axis=2)            # columns = edge vectors
        self.J = J
        self.detJ = J[:, 0, 0] * J[:, 1, 1] - J[:, 0, 1] * J[:, 1, 0]
        ad = np.abs(self.detJ)
        self.area = 0.5 * ad
        JtJ = np.einsum('cki,ckj->cij', J, J)
        # geometry the device needs: signed det J, G00, G01, G11
        self.geom = np.ascontiguousarray(np.stack([self.detJ, JtJ[:, 0, 0] / ad, JtJ[:, 0, 1] / ad, JtJ[:, 1, 1] / ad], axis=1))
        cnt = np.bincount(ce.ravel(), minlength=self.ne)
        self.boundary_edges = np.nonzero(cnt == 1)[0]
        r = self.ref
        # ---- P_k dofs
        k = kp
        nde, ndi = k - 1, (k - 1) * (k - 2) // 2
        self.ndp = r.ndp
        self.np_dofs = self.nv + self.ne * nde + self.nc * ndi
        cd = np.empty((self.nc, self.ndp), dtype=np.int64)
        cd[:, :3] = self.cells
        for e in range(3):
            for j in range(1, k):
                cd[:, 3 + e * nde + (j - 1)] = self.nv + ce[:, e] * nde + (j - 1)
        for i in range(ndi):
            cd[:, 3 + 3 * nde + i] = self.nv + self.ne * nde + np.arange(self.nc) * ndi + i
        self.cell_dofs_p = cd
        lam = np.stack([1 - r.p_nodes[:, 0] - r.p_nodes[:, 1], r.p_nodes[:, 0], r.p_nodes[:, 1]], axis=1)
        xy = np.einsum('nj,cjd->cnd', lam, V)
        self.p_coords = np.zeros((self.np_dofs, 2))
        self.p_coords[cd.ravel()] = xy.reshape(-1, 2)
        # ---- RT_k dofs
        k = kq
        self.ndq, self.n_int = r.ndq, r.n_int
        self.nq_dofs = self.ne * k + self.nc * self.n_int
        cq = np.empty((self.nc, self.ndq), dtype=np.int64)
        for e in range(3):
            for j in range(k):
                cq[:, e * k + j] = ce[:, e] * k + j
        for i in range(self.n_int):
            cq[:, 3 * k + i] = self.ne * k + np.arange(self.nc) * self.n_int + i
        self.cell_dofs_q = cq
        self._spec_maps = None

    # ---- element matrices (host copies of what the kernels apply; used for bounds, tests and the assembled fallbacks) ----------
    def local_Mq(self):
        g, r = self.geom, self.ref
        return g[:, 1, None, None] * r.A00[None] + g[:, 2, None, None] * r.A01s[None] + g[:, 3, None, None] * r.A11[None]

    def jacobi_bounds(self):
        """element-wise (Wathen) bounds of the Jacobi-scaled mass spectra"""
        def bounds(Mloc):
            dk = np.einsum('cii->ci', Mloc)
            ev = np.linalg.eigvalsh(Mloc / np.sqrt(dk[:, :, None] * dk[:, None, :]))
            return float(ev.min()), float(ev.max())
        return bounds(self.ref.Mp_hat[None]), bounds(self.local_Mq())

    # ---- boundary operators -------------------------------------------------------------------------------------------------
    def boundary_operators(self, btags_by_edge, left_weight=None, strong=False):
        """tags: dict boundary edge id -> tag (0 left, 1 top, 2 right, 3 general).  Returns N (zero-flux sides 1, 3; every side
        in the strong form) [n_q x n_p], b_L (weight g_L or 1) and b_L1 (weight 1) [n_q], B_top [n_p x n_p], B_right [n_q x n_q]."""
        r, kp, kq = self.ref, self.kp, self.kq
        be = self.boundary_edges
        tags = np.array([btags_by_edge.get(int(e), 4) for e in be])
        lo, hi = self.vertices[self.edges[be, 0]], self.vertices[self.edges[be, 1]]
        t = hi - lo
        L = np.hypot(t[:, 0], t[:, 1])
        # outward test: the vertex of the adjacent cell that is not on the edge lies on the inner side
        cell_of = np.full(self.ne, -1, dtype=np.int64)
        slot_of = np.zeros(self.ne, dtype=np.int64)
        for i in range(3):
            cell_of[self.cell_edges[:, i]] = np.arange(self.nc)
            slot_of[self.cell_edges[:, i]] = i
        opp = self.V[cell_of[be], slot_of[be]]
        rot = np.stack([t[:, 1], -t[:, 0]], axis=1)
        sigma = np.where(np.einsum('bd,bd->b', rot, 0.5 * (lo + hi) - opp) > 0, 1.0, -1.0)
        nde = kp - 1
        # P dofs on the edge: [lo vertex, hi vertex, interior points along lo -> hi]
        pd = np.concatenate([self.edges[be], self.nv + be[:, None] * nde + np.arange(nde)[None, :]], axis=1)      # [nb, kp+1]
        qd = be[:, None] * kq + np.arange(kq)[None, :]                                                              # [nb, kq]
        sel = np.ones(be.shape[0], dtype=bool) if strong else np.isin(tags, (1, 3))
        rows = np.repeat(qd[sel][:, :, None], kp + 1, axis=2)
        cols = np.repeat(pd[sel][:, None, :], kq, axis=1)
        vals = sigma[sel][:, None, None] * r.N_hat[None]
        N = sp.coo_matrix((vals.ravel(), (rows.ravel(), cols.ravel())), shape=(self.nq_dofs, self.np_dofs)).tocsr()
        bL, bL1 = np.zeros(self.nq_dofs), np.zeros(self.nq_dofs)
        left = tags == 0
        if left.any():
            one = sigma[left][:, None] * (r.edge_w @ r.trace_q)[None, :]
            bL1[qd[left].ravel()] = one.ravel()
            if left_weight is None:
                bL[qd[left].ravel()] = one.ravel()
            else:                                   # degree-8 interpolant of the datum on the facet (Expression degree=8, wave_2D.py:376-377)
                t9 = np.linspace(0.0, 1.0, 9)
                p9 = lo[left][:, None, :] + t9[None, :, None] * t[left][:, None, :]
                gq = left_weight(p9[..., 0], p9[..., 1]) @ _lagrange_matrix(t9, r.edge_s).T                         # [nl, nq]
                bL[qd[left].ravel()] = (sigma[left][:, None] * np.einsum('bq,qj,q->bj', gq, r.trace_q, r.edge_w)).ravel()
        top = tags == 1
        rows = np.repeat(pd[top][:, :, None], kp + 1, axis=2)
        cols = np.repeat(pd[top][:, None, :], kp + 1, axis=1)
        B_top = sp.coo_matrix(((L[top][:, None, None] * r.B1_hat[None]).ravel(), (rows.ravel(), cols.ravel())),
                              shape=(self.np_dofs, self.np_dofs)).tocsr()
        right = tags == 2
        dr = np.zeros(self.nq_dofs)
        dr[qd[right].ravel()] = ((2 * np.arange(kq) + 1)[None, :] / L[right][:, None]).ravel()
        B_right = sp.diags(dr).tocsr()
        return dict(N=N, b_L=bL, b_L1=bL1, B_top=B_top, B_right=B_right, sigma=sigma, tags=tags, edge_p_dofs=pd, edge_q_dofs=qd)

    def dirichlet_dofs(self, bops, left_weight=None):
        """strong Dirichlet rows (DirichletBC(U.sub(0), ...), wave_2D.py:413-422): P dofs on the left (value w g(t)) and right (0)"""
        tags, pd = bops['tags'], bops['edge_p_dofs']
        left = np.unique(pd[tags == 0].ravel())
        right = np.setdiff1d(np.unique(pd[tags == 2].ravel()), left)
        wl = np.ones(left.shape[0]) if left_weight is None else left_weight(self.p_coords[left, 0], self.p_coords[left, 1])
        return np.concatenate([left, right]), np.concatenate([wl, np.zeros(right.shape[0])])

    # ---- assembled operators (host; tests, bounds cross-checks) ----------------------------------------------------------
    def assemble(self, bops):
        r = self.ref

        def coo(rd, cdofs, loc, shape):
            rows = np.repeat(rd[:, :, None], cdofs.shape[1], axis=2)
            cols = np.repeat(cdofs[:, None, :], rd.shape[1], axis=1)
            return sp.coo_matrix((loc.ravel(), (rows.ravel(), cols.ravel())), shape=shape).tocsr()
        ad = np.abs(self.detJ)
        Mp = coo(self.cell_dofs_p, self.cell_dofs_p, ad[:, None, None] * r.Mp_hat[None], (self.np_dofs, self.np_dofs))
        Mq = coo(self.cell_dofs_q, self.cell_dofs_q, self.local_Mq(), (self.nq_dofs, self.nq_dofs))
        C = coo(self.cell_dofs_q, self.cell_dofs_p, np.sign(self.detJ)[:, None, None] * r.C_hat[None], (self.nq_dofs, self.np_dofs))
        D = (C - bops['N']).tocsr()
        return dict(M_p=Mp, M_q=Mq, C=C, D=D, b_L=bops['b_L'], b_L1=bops['b_L1'], B_top=bops['B_top'], B_right=bops['B_right'])

    # ---- the interior RT dofs of the specification (fem_ho.HighOrderSpaces / oracle.fem_ho) -----------------------------------
    def _spec(self):
        """per cell: spec interior moments (1/|K|) int_K q_c xi^a eta^b, (xi, eta) = (x - x_c) / sqrt(2|K|), of the local basis:
        L_K [n_int x ndq] = [X_K | Y_K]"""
        if self._spec_maps is None:
            r, k = self.ref, self.kq
            if self.n_int == 0:
                self._spec_maps = (np.zeros((self.nc, 0, self.ndq)), None)
                return self._spec_maps
            xh, yh = r.q_bary[:, 1], r.q_bary[:, 2]
            phi_hat, _ = r.tab_q(xh, yh)                                                       # [nq, ndq, 2]
            phi = np.einsum('cij,qnj->cqni', self.J, phi_hat) / self.detJ[:, None, None, None]   # physical values
            pts = np.einsum('qj,cjd->cqd', r.q_bary, self.V)
            xc, h = self.V.mean(axis=1), np.sqrt(2.0 * self.area)
            xi = (pts[..., 0] - xc[:, None, 0]) / h[:, None]
            eta = (pts[..., 1] - xc[:, None, 1]) / h[:, None]
            w = 2.0 * r.q_w                                                                      # weights summing to 1: (1/|K|) int_K
            Lm = np.zeros((self.nc, self.n_int, self.ndq))
            row = 0
            for bb in range(k - 1):
                for aa in range(k - 1 - bb):
                    m = xi ** aa * eta ** bb * w[None, :]
                    for comp in range(2):
                        Lm[:, row, :] = np.einsum('cqn,cq->cn', phi[..., comp], m)
                        row += 1
            Yinv = np.linalg.inv(Lm[:, :, 3 * k:])
            self._spec_maps = (Lm, Yinv)
        return self._spec_maps

    def to_spec(self, q_ref):
        """reference-basis dof vector -> dofs of the specification (edge moments unchanged)"""
        q = np.array(q_ref, dtype=np.float64)
        if self.n_int:
            Lm, _ = self._spec()
            q[self.cell_dofs_q[:, 3 * self.kq:]] = np.einsum('cin,cn->ci', Lm, q_ref[self.cell_dofs_q])
        return q

    def from_spec(self, q_spec):
        q = np.array(q_spec, dtype=np.float64)
        if self.n_int:
            Lm, Yinv = self._spec()
            ke = 3 * self.kq
            rhs = q_spec[self.cell_dofs_q[:, ke:]] - np.einsum('cin,cn->ci', Lm[:, :, :ke], q_spec[self.cell_dofs_q[:, :ke]])
            q[self.cell_dofs_q[:, ke:]] = np.einsum('cij,cj->ci', Yinv, rhs)
        return q

    def spec_transform(self):
        """sparse T with q_spec = T q_ref (tests)"""
        ke, ni = 3 * self.kq, self.n_int
        T = sp.identity(self.nq_dofs, format='lil')
        if ni:
            Lm, _ = self._spec()
            for c in range(self.nc):
                rows = self.cell_dofs_q[c, ke:]
                for i in range(ni):
                    T[rows[i], rows[i]] = 0.0
                    for n in range(self.ndq):
                        T[rows[i], self.cell_dofs_q[c, n]] = Lm[c, i, n]
        return T.tocsr()

    def interpolate_p(self, fun_xy):
        return fun_xy(self.p_coords[:, 0], self.p_coords[:, 1])

    def interpolate_q(self, fun_xy):
        """reference-basis dofs of a vector field (qx, qy) = fun_xy(x, y) through the functionals of the specification: edge
        moments int_e (q . rot t) L_j ds and interior moments (1/|K|) int_K q_c xi^a eta^b  (passivity initial condition,
        wave_2D.py:759-764)"""
        r, k = self.ref, self.kq
        q = np.zeros(self.nq_dofs)
        s, ws = r.edge_s, r.edge_w
        lo, hi = self.vertices[self.edges[:, 0]], self.vertices[self.edges[:, 1]]
        t = hi - lo
        pts = lo[:, None, :] + s[None, :, None] * t[:, None, :]
        fx, fy = fun_xy(pts[..., 0], pts[..., 1])
        fn = fx * t[:, None, 1] - fy * t[:, None, 0]
        for j in range(k):
            q[np.arange(self.ne) * k + j] = fn @ (ws * legendre01(j, s))
        if self.n_int:
            pts = np.einsum('qj,cjd->cqd', r.q_bary, self.V)
            f = np.stack(fun_xy(pts[..., 0], pts[..., 1]), axis=-1)
            xc, h = self.V.mean(axis=1), np.sqrt(2.0 * self.area)
            xi = (pts[..., 0] - xc[:, None, 0]) / h[:, None]
            eta = (pts[..., 1] - xc[:, None, 1]) / h[:, None]
            w = 2.0 * r.q_w
            row = 3 * k
            for bb in range(k - 1):
                for aa in range(k - 1 - bb):
                    m = xi ** aa * eta ** bb * w[None, :]
                    for comp in range(2):
                        q[self.cell_dofs_q[:, row]] = np.einsum('cq,cq->c', f[..., comp], m)
                        row += 1
        return self.from_spec(q)
