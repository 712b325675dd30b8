"""Oracle finite-element layer: topology + assembled sparse operators (numpy/scipy, fp64).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

This restates what dolfin/FFC does for the UFL forms of the reference
(``fenics_projects/wave_eqn/wave_2D.py:449-526`` wave forms, ``:528-577`` lumped rows,
``:600-641`` boundary energy forms) for the lowest-order pair used by every non-analytical
paper case: continuous P1 (``FiniteElement('P', triangle, 1)``, wave_2D.py:174,202) and
lowest-order Raviart-Thomas (``FiniteElement('RT', triangle, 1)``, wave_2D.py:175,203).

Everything is assembled by explicit quadrature of explicit basis functions, on purpose: the
CUDA path uses closed forms (geometry-free div/grad incidences, metric-based RT mass); an
oracle built the same way would not be an independent check.

Conventions (shared, by definition, with the product's GPU setup so that integer
indexing can be compared bit-exactly):
  * cells are re-oriented counter-clockwise;
  * local edge i of a cell is opposite local vertex i, i.e. joins (v[i+1], v[i+2]);
  * global edges are the unique (lo, hi) vertex pairs sorted lexicographically;
  * global edge normal = tangent (lo -> hi) rotated by -90 degrees, so for a CCW cell the
    local sign is +1 iff the cell traverses the edge lo -> hi;
  * RT dof = total flux through the edge in the direction of the global normal.
"""
import numpy as np
import scipy.sparse as sp

# boundary tags of wave_2D.py:266-280
TAG_LEFT, TAG_TOP, TAG_RIGHT, TAG_GENERAL, TAG_NONE = 0, 1, 2, 3, 4

# Dunavant 6-point rule, exact for degree 4 (weights sum to 1; multiply by |K|)
_A1, _B1, _W1 = 0.445948490915965, 0.108103018168070, 0.223381589678011
_A2, _B2, _W2 = 0.091576213509771, 0.816847572980459, 0.109951743655322
TRI_QP = np.array([[_A1, _A1, _B1], [_A1, _B1, _A1], [_B1, _A1, _A1],
                   [_A2, _A2, _B2], [_A2, _B2, _A2], [_B2, _A2, _A2]])
TRI_QW = np.array([_W1, _W1, _W1, _W2, _W2, _W2])
# Gauss-Legendre 5 points on [0,1]
_gx, _gw = np.polynomial.legendre.leggauss(5)
EDGE_QP = 0.5 * (_gx + 1.0)
EDGE_QW = 0.5 * _gw


class Topology:
    """Mesh connectivity derived from (vertices, cells) with the conventions above."""

    def __init__(self, vertices, cells):
        vertices = np.ascontiguousarray(vertices, dtype=np.float64)
        cells = np.array(cells, dtype=np.int64)
        v0, v1, v2 = (vertices[cells[:, i]] for i in range(3))
        area2 = (v1[:, 0] - v0[:, 0]) * (v2[:, 1] - v0[:, 1]) - (v1[:, 1] - v0[:, 1]) * (v2[:, 0] - v0[:, 0])
        flip = area2 < 0
        cells[flip, 1], cells[flip, 2] = cells[flip, 2].copy(), cells[flip, 1].copy()
        self.vertices = vertices
        self.cells = cells
        self.area = 0.5 * np.abs(area2)
        nv = vertices.shape[0]
        a = cells[:, [1, 2, 0]]
        b = cells[:, [2, 0, 1]]
        lo, hi = np.minimum(a, b), np.maximum(a, b)
        key = lo * nv + hi
        uniq, inv = np.unique(key.ravel(), return_inverse=True)
        self.edges = np.stack([uniq // nv, uniq % nv], axis=1)
        self.cell_edges = inv.reshape(-1, 3)
        self.cell_edge_sign = np.where(a < b, 1, -1).astype(np.int64)
        self.n_vertices, self.n_cells, self.n_edges = nv, cells.shape[0], self.edges.shape[0]
        count = np.bincount(self.cell_edges.ravel(), minlength=self.n_edges)
        self.boundary_edges = np.nonzero(count == 1)[0]
        # sign of the unique adjacent cell for boundary edges (+1: global normal is outward)
        bsign = np.zeros(self.n_edges, dtype=np.int64)
        np.add.at(bsign, self.cell_edges.ravel(), self.cell_edge_sign.ravel())
        self.boundary_sign = bsign[self.boundary_edges]
        e = vertices[self.edges[:, 1]] - vertices[self.edges[:, 0]]
        self.edge_length = np.hypot(e[:, 0], e[:, 1])


def tag_boundary(topo, domainShape, xLength, yLength):
    """Facet markers of wave_2D.py:226-280 (dolfin marks a facet when all its vertices and
    its midpoint satisfy ``inside``; later ``mark`` calls override earlier ones)."""
    def near(a, b):
        return np.abs(a - b) < 3.0e-16 * 1e2 + 1e-12  # dolfin DOLFIN_EPS-scale tolerance

    be = topo.boundary_edges
    pa, pb = topo.vertices[topo.edges[be, 0]], topo.vertices[topo.edges[be, 1]]
    pm = 0.5 * (pa + pb)
    xInputStart = 0.0 if domainShape == 'R' else -0.10

    def allpts(pred):
        return pred(pa) & pred(pb) & pred(pm)

    tags = np.full(be.shape[0], TAG_NONE, dtype=np.int64)
    tags[allpts(lambda p: near(p[:, 0], xInputStart))] = TAG_LEFT
    if domainShape == 'R':
        tags[allpts(lambda p: near(p[:, 1], yLength))] = TAG_TOP
    tags[allpts(lambda p: near(p[:, 0], xLength))] = TAG_RIGHT
    if domainShape == 'R':
        tags[allpts(lambda p: near(p[:, 1], 0.0))] = TAG_GENERAL
    else:
        tags[allpts(lambda p: near(p[:, 1], 0.0) | near(p[:, 1], xLength) |
                    near(p[:, 1], xLength / 2 - yLength / 2) |
                    near(p[:, 1], xLength / 2 + yLength / 2) | near(p[:, 0], 0.0))] = TAG_GENERAL
    return tags


def _coo(rows, cols, vals, shape):
    return sp.coo_matrix((np.ravel(vals), (np.ravel(rows), np.ravel(cols))), shape=shape).tocsr()


def _rt_basis_at(topo, lam):
    """phi_i(x_q) for every cell: array [n_cells, 3(i), n_q, 2].
    phi_i(x) = s_i (x - v_i) / (2|K|)  (flux-normalised lowest-order RT)."""
    V = topo.vertices[topo.cells]                      # [nc, 3, 2]
    xq = np.einsum('qj,cjd->cqd', lam, V)              # [nc, nq, 2]
    diff = xq[:, None, :, :] - V[:, :, None, :]        # [nc, 3, nq, 2]
    scale = topo.cell_edge_sign / (2.0 * topo.area[:, None])
    return diff * scale[:, :, None, None]


def assemble_wave_operators(topo, tags, neumann_tags=(TAG_TOP, TAG_GENERAL), left_weight=None):
    """Assemble the matrices the reference's forms reduce to.

    Returns a dict with
      M_p [Nv,Nv]  = (psi_i, psi_j)                      (p - p_m) v_p dx        wave_2D.py:452-518
      M_q [Ne,Ne]  = (phi_i, phi_j)                      dot(q - q_m, v_q) dx
      C   [Ne,Nv]  = (div phi_i, psi_j)                  div(v_q) p dx
      G   [Ne,Nv]  = (phi_i, grad psi_j)                 dot(grad(p), v_q) dx / dot(q, grad(v_p)) dx
      N   [Ne,Nv]  = <phi_i.n, psi_j> on `neumann_tags`  dot(v_q, n) p ds(1), ds(3)
      Nall[Ne,Nv]  = same on the whole boundary
      b_L [Ne]     = <phi_i.n, g_L> on tag 0             dot(v_q, n) pLeftBoundary ds(0)  (g_L = left_weight(y), default 1)
      b_R [Ne]     = <phi_i.n, 1> on tag 2
      B_top [Nv,Nv]   = <psi_i, psi_j> on tag 1          casimir impedance, wave_2D.py:330,476
      B_right [Ne,Ne] = <phi_i.n, phi_j.n> on tag 2      casimir impedance, wave_2D.py:350,482
      D = C - N
    """
    nc, nv, ne = topo.n_cells, topo.n_vertices, topo.n_edges
    lam = TRI_QP
    w = TRI_QW[None, :] * topo.area[:, None]                       # [nc, nq]
    psi = np.broadcast_to(lam.T[None, :, :], (nc, 3, lam.shape[0]))  # [nc, 3, nq]
    phi = _rt_basis_at(topo, lam)                                   # [nc, 3, nq, 2]
    V = topo.vertices[topo.cells]
    # grad lambda_j = rot(v_{j+2}-v_{j+1}) / (2|K|)  (CCW)
    ed = V[:, [2, 0, 1], :] - V[:, [1, 2, 0], :]                    # opposite-edge vectors
    grad = np.stack([-ed[:, :, 1], ed[:, :, 0]], axis=2) / (2.0 * topo.area[:, None, None])  # [nc,3,2]
    divphi = topo.cell_edge_sign / topo.area[:, None]               # [nc, 3]

    Mp_loc = np.einsum('ciq,cjq,cq->cij', psi, psi, w)
    Mq_loc = np.einsum('ciqd,cjqd,cq->cij', phi, phi, w)
    C_loc = np.einsum('ci,cjq,cq->cij', divphi, psi, w)
    G_loc = np.einsum('ciqd,cjd,cq->cij', phi, grad, w)

    cv, ce = topo.cells, topo.cell_edges
    rv = np.repeat(cv[:, :, None], 3, axis=2)
    cvj = np.repeat(cv[:, None, :], 3, axis=1)
    re = np.repeat(ce[:, :, None], 3, axis=2)
    cej = np.repeat(ce[:, None, :], 3, axis=1)
    ops = {
        'M_p': _coo(rv, cvj, Mp_loc, (nv, nv)),
        'M_q': _coo(re, cej, Mq_loc, (ne, ne)),
        'C': _coo(re, cvj, C_loc, (ne, nv)),
        'G': _coo(re, cvj, G_loc, (ne, nv)),
    }
    # ---- boundary (exterior facet) terms: on boundary edge e, phi_e.n_out = bsign/|e|, other RT
    # functions have zero normal flux; psi restricted to the edge is linear.
    be, bs, L = topo.boundary_edges, topo.boundary_sign.astype(np.float64), topo.edge_length[topo.boundary_edges]
    va, vb = topo.edges[be, 0], topo.edges[be, 1]
    s = EDGE_QP
    psi_a, psi_b = 1.0 - s, s
    flux = bs / L                                                   # phi_e . n_out on e
    Na = flux * L * np.sum(EDGE_QW * psi_a)                         # = bs/2
    Nb = flux * L * np.sum(EDGE_QW * psi_b)

    def Nmat(mask):
        return _coo(np.concatenate([be[mask], be[mask]]), np.concatenate([va[mask], vb[mask]]),
                    np.concatenate([Na[mask], Nb[mask]]), (ne, nv))
    ops['N'] = Nmat(np.isin(tags, neumann_tags))
    ops['Nall'] = Nmat(np.ones_like(tags, dtype=bool))
    ops['D'] = (ops['C'] - ops['N']).tocsr()
    ops['Dall'] = (ops['C'] - ops['Nall']).tocsr()

    bL = np.zeros(ne)
    left = tags == TAG_LEFT
    if left_weight is None:
        bL[be[left]] = flux[left] * L[left]
    else:
        # the reference's left datum is a degree=8 Expression (wave_2D.py:376-377): dolfin
        # interpolates it into P8 on each cell; its trace on an edge is the degree-8 Lagrange
        # interpolant at 9 equispaced points, integrated exactly = closed Newton-Cotes(9).
        pa, pb = topo.vertices[va[left]], topo.vertices[vb[left]]
        bL[be[left]] = flux[left] * L[left] * newton_cotes9_mean(left_weight, pa, pb)
    ops['b_L'] = bL
    bR = np.zeros(ne)
    right = tags == TAG_RIGHT
    bR[be[right]] = flux[right] * L[right]
    ops['b_R'] = bR
    top = tags == TAG_TOP
    maa = L * np.sum(EDGE_QW * psi_a * psi_a)
    mab = L * np.sum(EDGE_QW * psi_a * psi_b)
    ops['B_top'] = _coo(np.concatenate([va[top], va[top], vb[top], vb[top]]),
                        np.concatenate([va[top], vb[top], va[top], vb[top]]),
                        np.concatenate([maa[top], mab[top], mab[top], maa[top]]), (nv, nv))
    ops['B_right'] = _coo(be[right], be[right], (flux * flux * L)[right], (ne, ne))
    return ops


# closed Newton-Cotes weights for 9 equispaced points on [0,1] (degree-8 interpolatory rule)
_NC9 = np.array([989.0, 5888.0, -928.0, 10496.0, -4540.0, 10496.0, -928.0, 5888.0, 989.0]) / 28350.0


def newton_cotes9_mean(fun_of_xy, pa, pb):
    """(1/|e|) * integral over edge [pa,pb] of the degree-8 equispaced interpolant of fun."""
    t = np.linspace(0.0, 1.0, 9)
    pts = pa[:, None, :] * (1.0 - t)[None, :, None] + pb[:, None, :] * t[None, :, None]
    vals = fun_of_xy(pts[..., 0], pts[..., 1])
    return vals @ _NC9


# ----------------------------------------------------------------------------------------------
# Lagrange P_k nodal basis on the reference triangle (used by the analytical diagnostics:
# dolfin errornorm interpolates into P_{k+3}, wave_2D.py:750,956)
# ----------------------------------------------------------------------------------------------
def lagrange_nodes(k):
    pts = []
    for j in range(k + 1):
        for i in range(k + 1 - j):
            pts.append((i / k, j / k))
    return np.array(pts)          # (x, y) reference coordinates; lambda = (1-x-y, x, y)


def _monomials(k, x, y):
    return np.stack([x ** a * y ** b for b in range(k + 1) for a in range(k + 1 - b)], axis=-1)


def lagrange_tabulate(k, x, y):
    """Values of the P_k nodal basis (equispaced nodes) at reference points: [n_pts, n_nodes]."""
    nodes = lagrange_nodes(k)
    Vn = _monomials(k, nodes[:, 0], nodes[:, 1])
    return np.linalg.solve(Vn.T, _monomials(k, x, y).T).T


def high_order_tri_rule(n=12):
    """Collapsed Gauss-Legendre (Duffy) rule with n*n points, exact to degree 2n-2 on the triangle."""
    gx, gw = np.polynomial.legendre.leggauss(n)
    u, wu = 0.5 * (gx + 1), 0.5 * gw
    U, Vv = np.meshgrid(u, u, indexing='ij')
    W = np.outer(wu, wu) * (1 - U)
    x = U.ravel()
    y = (Vv * (1 - U)).ravel()
    return x, y, 2.0 * W.ravel()   # weights sum to 1 (times |K| gives the integral)
