"""Thin object wrappers over the C ABI handles (include/phwave.h)."""
import ctypes

import numpy as np

from . import _lib

TAG_LEFT, TAG_TOP, TAG_RIGHT, TAG_GENERAL, TAG_NONE = 0, 1, 2, 3, 4


class Mesh:
    """Connectivity + patch layout of a triangle mesh (phw_mesh_*; host only, no GPU needed)."""

    def __init__(self, vertices, cells, patch_cells=2048, vertex_external=None, edge_external=None, cell_group=None, n_groups=1,
                 cell_owned=None):
        self.lib = _lib.load()
        self.vertices = np.ascontiguousarray(vertices, dtype=np.float64)
        self.cells = np.ascontiguousarray(cells, dtype=np.int32)
        if self.vertices.ndim != 2 or self.vertices.shape[1] != 2 or self.cells.ndim != 2 or self.cells.shape[1] != 3:
            raise ValueError('vertices must be [Nv,2], cells must be [Nc,3]')
        h = ctypes.c_void_p()
        self.n_groups = int(n_groups) if cell_group is not None else 1
        if cell_group is not None:   # batched sweep: disjoint copies of one domain, one case per copy
            cg = np.ascontiguousarray(cell_group, dtype=np.int32)
            _lib.check(self.lib.phw_mesh_create_grouped(self.vertices.shape[0], self.vertices.ctypes.data_as(_lib.c_f64p),
                                                        self.cells.shape[0], self.cells.ctypes.data_as(_lib.c_i32p),
                                                        int(patch_cells), cg.ctypes.data_as(_lib.c_i32p), self.n_groups, ctypes.byref(h)))
        elif vertex_external is None:
            _lib.check(self.lib.phw_mesh_create(self.vertices.shape[0], self.vertices.ctypes.data_as(_lib.c_f64p),
                                                self.cells.shape[0], self.cells.ctypes.data_as(_lib.c_i32p),
                                                int(patch_cells), ctypes.byref(h)))
        else:   # local mesh of one rank of a partitioned run
            ve = np.ascontiguousarray(vertex_external, dtype=np.uint8)
            ee = np.ascontiguousarray(edge_external, dtype=np.uint8)
            if ve.shape[0] != self.vertices.shape[0]:
                raise ValueError('one external flag per vertex expected')
            _lib.check(self.lib.phw_mesh_create_partitioned(
                self.vertices.shape[0], self.vertices.ctypes.data_as(_lib.c_f64p), self.cells.shape[0],
                self.cells.ctypes.data_as(_lib.c_i32p), int(patch_cells), ve.ctypes.data, ee.shape[0], ee.ctypes.data,
                ctypes.byref(h)))
        self.handle = h
        if cell_owned is not None:   # partitioned run: which cells of the local mesh are this rank's (cell-based reductions)
            co = np.ascontiguousarray(cell_owned, dtype=np.uint8)
            if co.shape[0] != self.cells.shape[0]:
                raise ValueError('one ownership flag per cell expected')
            _lib.check(self.lib.phw_mesh_set_cell_owned(h, co.ctypes.data))
        s = [ctypes.c_int64() for _ in range(5)]
        _lib.check(self.lib.phw_mesh_sizes(h, *[ctypes.byref(x) for x in s]))
        self.n_vertices, self.n_cells, self.n_edges, self.n_boundary, self.n_patches = [int(x.value) for x in s]
        self._edges = None
        self.tags = None

    def edges(self):
        """(edges [Ne,2], cell_edges [Nc,3], cell_edge_sign [Nc,3]) in the user numbering."""
        if self._edges is None:
            e = np.empty((self.n_edges, 2), dtype=np.int32)
            ce = np.empty((self.n_cells, 3), dtype=np.int32)
            cs = np.empty((self.n_cells, 3), dtype=np.int8)
            _lib.check(self.lib.phw_mesh_get_edges(self.handle, e.ctypes.data_as(_lib.c_i32p),
                                                   ce.ctypes.data_as(_lib.c_i32p), cs.ctypes.data_as(_lib.c_i8p)))
            self._edges = (e, ce, cs)
        return self._edges

    def boundary(self):
        ids = np.empty(self.n_boundary, dtype=np.int32)
        sg = np.empty(self.n_boundary, dtype=np.int8)
        _lib.check(self.lib.phw_mesh_get_boundary(self.handle, ids.ctypes.data_as(_lib.c_i32p), sg.ctypes.data_as(_lib.c_i8p)))
        return ids, sg

    def set_boundary(self, tags, port_weight=None):
        tags = np.ascontiguousarray(tags, dtype=np.int32)
        if tags.shape[0] != self.n_boundary:
            raise ValueError('one tag per boundary edge expected')
        w = None
        if port_weight is not None:
            port_weight = np.ascontiguousarray(port_weight, dtype=np.float64)
            w = port_weight.ctypes.data_as(_lib.c_f64p)
        _lib.check(self.lib.phw_mesh_set_boundary(self.handle, tags.ctypes.data_as(_lib.c_i32p), w))
        self.tags = tags

    def set_dirichlet(self, vertex_ids, weights):
        vertex_ids = np.ascontiguousarray(vertex_ids, dtype=np.int32)
        weights = np.ascontiguousarray(weights, dtype=np.float64)
        _lib.check(self.lib.phw_mesh_set_dirichlet(self.handle, vertex_ids.shape[0], vertex_ids.ctypes.data_as(_lib.c_i32p),
                                                   weights.ctypes.data_as(_lib.c_f64p)))

    def numbering(self):
        v = np.empty(self.n_vertices, dtype=np.int32)
        e = np.empty(self.n_edges, dtype=np.int32)
        c = np.empty(self.n_cells, dtype=np.int32)
        _lib.check(self.lib.phw_mesh_get_numbering(self.handle, v.ctypes.data_as(_lib.c_i32p),
                                                   e.ctypes.data_as(_lib.c_i32p), c.ctypes.data_as(_lib.c_i32p)))
        return v, e, c

    def self_check(self):
        """structural invariants of the layout (host-only); raises PhwError with the first violation"""
        _lib.check(self.lib.phw_mesh_self_check(self.handle))

    def patch_counts(self):
        c = np.empty((self.n_patches, 6), dtype=np.int32)
        _lib.check(self.lib.phw_mesh_get_patch_counts(self.handle, c.ctypes.data_as(_lib.c_i32p)))
        return c

    def pipe_smem(self):
        """dynamic shared memory (bytes) of the bulk-copy pipelined kernels for this layout:
        (M_p solve, M_q solve, energy reduction, edge-row rhs, vertex-row rhs, reserved) - flags bits 9, 8, 11, 12, 12"""
        b = (ctypes.c_int64 * 6)()
        _lib.check(self.lib.phw_mesh_pipe_smem(self.handle, b))
        return tuple(int(x) for x in b)

    def close(self):
        if getattr(self, 'handle', None):
            self.lib.phw_mesh_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def facet_tags(pa, pb, domainShape, xLength, yLength):
    """Facet markers of the reference (wave_2D.py:226-280) for boundary facets with end points pa, pb [n, 2]: 0 left/port, 1 top
    (rectangle only), 2 right, 3 bottom/general, 4 unmarked.  dolfin marks a facet when its vertices and midpoint satisfy ``inside``;
    later markers override earlier ones (left, top, right, general)."""
    pm = 0.5 * (pa + pb)
    tol = 1e-12
    xInputStart = 0.0 if domainShape == 'R' else -0.10

    def on(pred):
        return pred(pa) & pred(pb) & pred(pm)

    tags = np.full(pa.shape[0], TAG_NONE, dtype=np.int32)
    tags[on(lambda p: np.abs(p[:, 0] - xInputStart) < tol)] = TAG_LEFT
    if domainShape == 'R':
        tags[on(lambda p: np.abs(p[:, 1] - yLength) < tol)] = TAG_TOP
    tags[on(lambda p: np.abs(p[:, 0] - xLength) < tol)] = TAG_RIGHT
    if domainShape == 'R':
        tags[on(lambda p: np.abs(p[:, 1]) < tol)] = TAG_GENERAL
    else:
        ylo, yhi = xLength / 2 - yLength / 2, xLength / 2 + yLength / 2
        tags[on(lambda p: (np.abs(p[:, 1]) < tol) | (np.abs(p[:, 1] - xLength) < tol) | (np.abs(p[:, 1] - ylo) < tol) |
                (np.abs(p[:, 1] - yhi) < tol) | (np.abs(p[:, 0]) < tol))] = TAG_GENERAL
    return tags


def tag_boundary_edges(mesh, domainShape, xLength, yLength):
    """facet_tags for the boundary edges of a libphwave Mesh (order of mesh.boundary())"""
    ids, _ = mesh.boundary()
    edges = mesh.edges()[0]
    return facet_tags(mesh.vertices[edges[ids, 0]], mesh.vertices[edges[ids, 1]], domainShape, xLength, yLength)


def tag_space_boundary(spaces, domainShape, xLength, yLength):
    """{boundary edge id: tag} for the assembled / reference-element spaces of fem_ho.py / fem_ref.py"""
    be = spaces.boundary_edges
    tags = facet_tags(spaces.vertices[spaces.edges[be, 0]], spaces.vertices[spaces.edges[be, 1]], domainShape, xLength, yLength)
    return {int(e): int(t) for e, t in zip(be, tags)}


# closed 9-point Newton-Cotes weights on [0,1]: exact integral of the degree-8 equispaced interpolant, which
# is what dolfin integrates for a ``degree=8`` Expression restricted to a facet (wave_2D.py:376-377)
_NC9 = np.array([989.0, 5888.0, -928.0, 10496.0, -4540.0, 10496.0, -928.0, 5888.0, 989.0]) / 28350.0


def port_weights(mesh, tags, fun_xy):
    """(1/|e|) int_e g_L ds for the left-port edges (1 elsewhere)."""
    ids, _ = mesh.boundary()
    edges = mesh.edges()[0]
    w = np.ones(ids.shape[0], dtype=np.float64)
    left = tags == TAG_LEFT
    pa = mesh.vertices[edges[ids[left], 0]]
    pb = mesh.vertices[edges[ids[left], 1]]
    t = np.linspace(0.0, 1.0, 9)
    pts = pa[:, None, :] * (1 - t)[None, :, None] + pb[:, None, :] * t[None, :, None]
    w[left] = fun_xy(pts[..., 0], pts[..., 1]) @ _NC9
    return w


class Solver:
    """Device-resident solver (phw_solver_*).  Raises if no CUDA device / library is available."""

    def __init__(self, mesh, params, device=0, stream=None):
        self.lib = _lib.load()
        self.mesh = mesh
        self.params = params
        h = ctypes.c_void_p()
        _lib.check(self.lib.phw_solver_create(mesh.handle, ctypes.byref(params), int(device),
                                              ctypes.c_void_p(stream) if stream else None, ctypes.byref(h)))
        self.handle = h
        self.ncol = 12 if params.analytical else 8

    def set_state(self, p=None, q=None, x3=None):
        keep = [None if a is None else (a if not isinstance(a, np.ndarray) else np.ascontiguousarray(a, dtype=np.float64)) for a in (p, q, x3)]
        _lib.check(self.lib.phw_set_state(self.handle, *[_lib.ptr(a) for a in keep]))

    def get_state(self):
        p = np.empty(self.mesh.n_vertices)
        q = np.empty(self.mesh.n_edges)
        x = np.empty(3)
        _lib.check(self.lib.phw_get_state(self.handle, p.ctypes.data, q.ctypes.data, x.ctypes.data))
        return p, q, x

    def set_group_params(self, K_wave, rho, lumped6):
        K_wave = np.ascontiguousarray(K_wave, dtype=np.float64)
        rho = np.ascontiguousarray(rho, dtype=np.float64)
        lumped6 = np.ascontiguousarray(lumped6, dtype=np.float64)
        _lib.check(self.lib.phw_solver_set_group_params(self.handle, K_wave.shape[0], K_wave.ctypes.data_as(_lib.c_f64p),
                                                        rho.ctypes.data_as(_lib.c_f64p), lumped6.ctypes.data_as(_lib.c_f64p)))

    def group_states(self):
        x = np.empty((self.mesh.n_groups, 3))
        _lib.check(self.lib.phw_get_group_states(self.handle, x.ctypes.data_as(_lib.c_f64p)))
        return x

    def run(self, n_steps, out=None):
        if out is None:
            G = getattr(self.mesh, 'n_groups', 1)
            out = np.zeros((n_steps, self.ncol)) if G == 1 else np.zeros((G, n_steps, self.ncol))
        _lib.check(self.lib.phw_run(self.handle, int(n_steps), _lib.ptr(out)))
        return out

    def run_with_snapshots(self, n_steps, every, n_values=None):
        """n_steps steps; the p field after every `every` steps (user numbering) -> (H_rows, snapshots [n_steps // every, n]).
        Device-side snapshot ring + second stream (phw_set_snapshots): replaces xdmfFile_p.write(out_p, t), wave_2D.py:865-866."""
        n = int(n_values if n_values is not None else self.mesh.n_vertices)
        snaps = np.zeros((int(n_steps) // int(every), n))
        _lib.check(self.lib.phw_set_snapshots(self.handle, int(every), snaps.ctypes.data, snaps.shape[0]))
        try:
            H = self.run(n_steps)
        finally:
            _lib.check(self.lib.phw_set_snapshots(self.handle, 0, None, 0))
        return H, snaps

    def stats(self):
        st = _lib.PhwStats()
        _lib.check(self.lib.phw_get_stats(self.handle, ctypes.byref(st)))
        return st.as_dict()

    def set_time(self, t):
        _lib.check(self.lib.phw_set_time(self.handle, float(t)))

    def stream_probe(self, threads=256, ctas_per_sm=8, reps=10):
        """(ms, bytes) of a pass that only streams the data of one ELL M_p iteration (diagnostics)"""
        ms, nb = ctypes.c_double(0), ctypes.c_double(0)
        _lib.check(self.lib.phw_stream_probe(self.handle, int(threads), int(ctas_per_sm), int(reps), ctypes.byref(ms), ctypes.byref(nb)))
        return ms.value, nb.value

    def residual_history(self, which, n=48):
        """relative residuals of the iterations of the most recent solve of kind `which` (diagnostics)"""
        out = np.zeros(n)
        nit = ctypes.c_int32(0)
        _lib.check(self.lib.phw_residual_history(self.handle, int(which), int(n), out.ctypes.data_as(_lib.c_f64p), ctypes.byref(nit)))
        return out[:min(n, nit.value)]

    def set_H_init(self, v):
        _lib.check(self.lib.phw_set_H_init(self.handle, float(v)))

    def energy_parts(self):
        out = (ctypes.c_double * 2)()
        _lib.check(self.lib.phw_energy(self.handle, out))
        return float(out[0]), float(out[1])

    def port_flux(self):
        out = ctypes.c_double()
        _lib.check(self.lib.phw_port_flux(self.handle, ctypes.byref(out)))
        return float(out.value)

    def apply_D(self, p):
        p = np.ascontiguousarray(p, dtype=np.float64)
        y = np.empty(self.mesh.n_edges)
        _lib.check(self.lib.phw_apply_D(self.handle, p.ctypes.data, y.ctypes.data))
        return y

    def apply_DT(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64)
        y = np.empty(self.mesh.n_vertices)
        _lib.check(self.lib.phw_apply_DT(self.handle, q.ctypes.data, y.ctypes.data))
        return y

    def mass_matvec(self, which, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty_like(x)
        _lib.check(self.lib.phw_mass_matvec(self.handle, int(which), x.ctypes.data, y.ctypes.data))
        return y

    def mass_solve(self, which, b):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.empty_like(b)
        it = ctypes.c_int32()
        _lib.check(self.lib.phw_mass_solve(self.handle, int(which), b.ctypes.data, x.ctypes.data, ctypes.byref(it)))
        return x, int(it.value)

    def close(self):
        if getattr(self, 'handle', None):
            self.lib.phw_solver_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def setup_peer_exchange(solver, part, dist, rank, world):
    """Wire the peer-memory halo exchange of a partitioned run (one process per GPU).

    `part` is a partition.LocalPart, `dist` an initialised torch.distributed (any backend; only object
    collectives are used here - the data path itself is peer-memory stores issued by the solve kernel)."""
    from . import partition as pt
    lib = solver.lib
    # 1. identical Chebyshev interval on every rank
    st = solver.stats()
    lam = [None] * world
    dist.all_gather_object(lam, (st['lam_min_q'], st['lam_max_q']))
    _lib.check(lib.phw_set_cheb_bounds(solver.handle, min(x[0] for x in lam), max(x[1] for x in lam)))
    # 2. where my external dofs live in my vectors, and which global dofs they are
    need = {}
    for nb, lists in part.recv.items():
        need[nb] = {}
        for kind, idx in lists.items():
            idx = np.ascontiguousarray(idx, dtype=np.int32)
            internal = np.empty_like(idx)
            _lib.check(lib.phw_comm_internal_index(solver.handle, 0 if kind == 'v' else 1, idx.shape[0],
                                                   idx.ctypes.data_as(_lib.c_i32p), internal.ctypes.data_as(_lib.c_i32p)))
            gid = (part.vertex_gid if kind == 'v' else part.edge_gid)[idx]
            need[nb][kind] = (gid, internal)
    allneed = [None] * world
    dist.all_gather_object(allneed, need)
    # 3. arena handles and local sizes
    handle = (ctypes.c_char * 64)()
    nvl, nel = ctypes.c_int64(), ctypes.c_int64()
    _lib.check(lib.phw_comm_arena_handle(solver.handle, handle, ctypes.byref(nvl), ctypes.byref(nel)))
    allh = [None] * world
    dist.all_gather_object(allh, (bytes(handle), int(nvl.value), int(nel.value)))
    handles = b''.join(x[0] for x in allh)
    peer_nv = np.array([x[1] for x in allh], dtype=np.int64)
    peer_ne = np.array([x[2] for x in allh], dtype=np.int64)
    # 4. send lists: what each other rank asked me for
    nbs = sorted(r for r in range(world) if r != rank and (rank in allneed[r] or r in part.recv))
    keep = []
    n_send = {'v': [], 'e': []}
    loc_ptrs = {'v': [], 'e': []}
    rem_ptrs = {'v': [], 'e': []}
    for nb in nbs:
        asked = allneed[nb].get(rank, {})
        for kind in ('v', 'e'):
            if kind in asked:
                gids, remote = asked[kind]
                loc = np.ascontiguousarray(pt.needed_by(gids, part, kind), dtype=np.int32)
                rem = np.ascontiguousarray(remote, dtype=np.int32)
            else:
                loc = np.zeros(0, dtype=np.int32)
                rem = np.zeros(0, dtype=np.int32)
            keep += [loc, rem]
            n_send[kind].append(loc.shape[0])
            loc_ptrs[kind].append(loc.ctypes.data_as(_lib.c_i32p))
            rem_ptrs[kind].append(rem.ctypes.data_as(_lib.c_i32p))
    n = len(nbs)
    PtrArr = _lib.c_i32p * max(n, 1)
    nb_arr = np.array(nbs, dtype=np.int32) if n else np.zeros(1, dtype=np.int32)
    nsv = np.array(n_send['v'] or [0], dtype=np.int64)
    nse = np.array(n_send['e'] or [0], dtype=np.int64)
    _lib.check(lib.phw_comm_setup(
        solver.handle, rank, world, handles, peer_nv.ctypes.data_as(_lib.c_i64p), peer_ne.ctypes.data_as(_lib.c_i64p),
        n, nb_arr.ctypes.data_as(_lib.c_i32p),
        nsv.ctypes.data_as(_lib.c_i64p), PtrArr(*loc_ptrs['v']) if n else PtrArr(), PtrArr(*rem_ptrs['v']) if n else PtrArr(),
        nse.ctypes.data_as(_lib.c_i64p), PtrArr(*loc_ptrs['e']) if n else PtrArr(), PtrArr(*rem_ptrs['e']) if n else PtrArr()))
    dist.barrier()
    info = dict(neighbours=nbs, send_v=n_send['v'], send_e=n_send['e'])
    if solver.params.scheme in (_lib.SCHEMES['IE'], _lib.SCHEMES['SM']):
        # implicit schemes: the coupled block solve needs the skew bound of the GLOBAL operator and (lumped model) a bordered column
        # solved with the exchange in place
        send_user = {'v': {}, 'e': {}}
        for nb in nbs:
            for kind in ('v', 'e'):
                asked = allneed[nb].get(rank, {})
                if kind in asked:
                    send_user[kind][nb] = pt.needed_by(asked[kind][0], part, kind)
        sigma = distributed_skew_estimate(solver, part, dist, rank, world, send_user)
        _lib.check(lib.phw_comm_finish_implicit(solver.handle, float(sigma)))
        dist.barrier()
        info['skew_sigma'] = float(sigma)
    return info


def distributed_skew_estimate(solver, part, dist, rank, world, send_user, iters=40):
    """Largest singular value of P_q^-1/2 D P_p^-1/2 of the global operator by power iteration on D^T P_q^-1 D (the same iteration as
    estimate_skew in phwave.cu, which sees one rank's operator only): D and D^T are applied on the device to the owned rows of
    every rank, the halo values and the norms travel through torch.distributed object collectives (set-up only).  The start vector
    depends on global vertex ids only, and sums run in rank order, so every rank ends with the same bits."""
    ov, oe = part.vertex_owner == rank, part.edge_owner == rank

    def halo(vec, kind):
        out = {int(nb): vec[idx] for nb, idx in send_user[kind].items()}
        allv = [None] * world
        dist.all_gather_object(allv, out)
        for nb, lists in part.recv.items():
            if kind in lists:
                vec[lists[kind]] = allv[nb][rank]

    def gsum(x):
        allx = [None] * world
        dist.all_gather_object(allx, float(x))
        return float(sum(allx))

    dp, dq = np.empty(ov.shape[0]), np.empty(oe.shape[0])
    _lib.check(solver.lib.phw_get_jacobi(solver.handle, _lib.PHW_P, dp.ctypes.data))
    _lib.check(solver.lib.phw_get_jacobi(solver.handle, _lib.PHW_Q, dq.ctypes.data))
    halo(dp, 'v')
    halo(dq, 'e')
    sp = np.sqrt(dp)
    gid = np.asarray(part.vertex_gid, dtype=np.uint64)
    h = ((gid * np.uint64(2654435761)) % np.uint64(2000001)).astype(np.float64) / 1e6 - 1.0
    lam = 0.0
    for _ in range(iters):
        y = solver.apply_D(h * sp) * dq
        y[~oe] = 0.0
        halo(y, 'e')
        w = solver.apply_DT(y) * sp
        w[~ov] = 0.0
        nrm = np.sqrt(gsum(w[ov] @ w[ov]))
        if nrm == 0.0:
            break
        lam = nrm / np.sqrt(gsum(h[ov] @ h[ov]))
        h = w / nrm
        halo(h, 'v')
    return np.sqrt(lam)


class _DofCounts:
    """minimal stand-in for Mesh in solvers that are not tied to a libphwave mesh handle"""

    def __init__(self, n_p, n_q):
        self.n_vertices, self.n_edges, self.n_groups = int(n_p), int(n_q), 1
        self.handle = None


class HighOrderSolver(Solver):
    """Solver for the P_k x RT_k pairs, k <= 3, on host-assembled CSR operators (phw_ho_create)."""

    def __init__(self, spaces, ops, params, bounds_p, bounds_q, device=0):
        self.lib = _lib.load()
        self.spaces = spaces
        self.mesh = _DofCounts(spaces.np_dofs, spaces.nq_dofs)
        self.params = params
        keep = []

        def csr(M):
            M = M.tocsr()
            M.sort_indices()
            ip = np.ascontiguousarray(M.indptr, dtype=np.int32)
            ix = np.ascontiguousarray(M.indices, dtype=np.int32)
            v = np.ascontiguousarray(M.data, dtype=np.float64)
            keep.extend([ip, ix, v])
            return ip.ctypes.data_as(_lib.c_i32p), ix.ctypes.data_as(_lib.c_i32p), v.ctypes.data_as(_lib.c_f64p)
        bL = np.ascontiguousarray(ops['b_L'], dtype=np.float64)
        h = ctypes.c_void_p()
        _lib.check(self.lib.phw_ho_create(ctypes.byref(params), int(device), None, spaces.np_dofs, spaces.nq_dofs,
                                          *csr(ops['M_p']), *csr(ops['M_q']), *csr(ops['D']), *csr(ops['D'].T),
                                          bL.ctypes.data_as(_lib.c_f64p), bounds_p[0], bounds_p[1], bounds_q[0], bounds_q[1],
                                          ctypes.byref(h)))
        self.handle = h
        self.ncol = 12 if params.analytical else 8


class BatchedSweepSolver(Solver):
    """n_cases independent cases on ONE set of assembled operators, state vectors [n_dof][n_cases] with the case index fastest
    (phw_sweep_create, csrc/kernels_sweep.cuh): the batched parameter sweep of BASELINE configs[4]."""

    def __init__(self, spaces, ops, params, bounds_p, bounds_q, n_cases, device=0):
        self.lib = _lib.load()
        self.spaces = spaces
        self.n_cases = int(n_cases)
        self.mesh = _DofCounts(spaces.np_dofs * self.n_cases, spaces.nq_dofs * self.n_cases)
        self.mesh.n_groups = self.n_cases
        self.params = params
        keep = []

        def csr(M):
            M = M.tocsr()
            M.sort_indices()
            ip = np.ascontiguousarray(M.indptr, dtype=np.int32)
            ix = np.ascontiguousarray(M.indices, dtype=np.int32)
            v = np.ascontiguousarray(M.data, dtype=np.float64)
            keep.extend([ip, ix, v])
            return ip.ctypes.data_as(_lib.c_i32p), ix.ctypes.data_as(_lib.c_i32p), v.ctypes.data_as(_lib.c_f64p)
        bL = np.ascontiguousarray(ops['b_L'], dtype=np.float64)
        h = ctypes.c_void_p()
        _lib.check(self.lib.phw_sweep_create(ctypes.byref(params), int(device), None, self.n_cases, spaces.np_dofs, spaces.nq_dofs,
                                             *csr(ops['M_p']), *csr(ops['M_q']), *csr(ops['D']), *csr(ops['D'].T),
                                             bL.ctypes.data_as(_lib.c_f64p), bounds_p[0], bounds_p[1], bounds_q[0], bounds_q[1],
                                             ctypes.byref(h)))
        self.handle = h
        self.ncol = 8

    def get_case_states(self):
        """(p [n_cases, n_p], q [n_cases, n_q], x [n_cases, 3]) - the device layout is [dof][case]"""
        p, q, _ = self.get_state()
        return (p.reshape(self.spaces.np_dofs, self.n_cases).T.copy(), q.reshape(self.spaces.nq_dofs, self.n_cases).T.copy(), self.group_states())


class MatrixFreeHighOrderSolver(Solver):
    """Solver for the P_k x RT_k pairs, k <= 3, with the matrix-free kernels of csrc/kernels_ho.cuh (phw_ho_create_mf): the device
    holds 3 vertex ids + 3 edge ids + 4 doubles per cell and the reference tabulations of fem_ref.RefSpaces; nothing is assembled.
    State vectors at this interface are in the dofs of the specification (fem_ho / oracle.fem_ho); the interior RT dofs are converted
    to / from the reference basis the device works in."""

    def __init__(self, spaces, bops, params, bounds_p, bounds_q, device=0):
        self.lib = _lib.load()
        self.spaces = spaces
        self.mesh = _DofCounts(spaces.np_dofs, spaces.nq_dofs)
        self.params = params
        r = spaces.ref
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)

        def compact(M):
            """rows of a sparse matrix that hold entries -> (n_rows, row ids, indptr, idx, val)"""
            M = M.tocsr()
            M.eliminate_zeros()
            M.sort_indices()
            rows = np.nonzero(np.diff(M.indptr))[0]
            ip = np.concatenate([[0], np.cumsum(np.diff(M.indptr)[rows])])
            return int(rows.shape[0]), i32(rows), i32(ip), i32(M.indices), f64(M.data)
        nN, N_row, N_ip, N_ix, N_va = compact(bops['N'])
        nT, T_row, T_ip, T_ix, T_va = compact(bops['N'].T)
        cv, ce, geom = i32(spaces.cells), i32(spaces.cell_edges), f64(spaces.geom)
        ref = [f64(a) for a in (r.Mp_hat, r.C_hat, r.A00, r.A01s, r.A11)]
        bL = f64(bops['b_L'])
        P = lambda a: a.ctypes.data_as(_lib.c_i32p if a.dtype == np.int32 else _lib.c_f64p)
        h = ctypes.c_void_p()
        _lib.check(self.lib.phw_ho_create_mf(ctypes.byref(params), int(device), None, int(spaces.kp), int(spaces.kq),
                                             spaces.nv, spaces.ne, spaces.nc, P(cv), P(ce), P(geom), *[P(a) for a in ref],
                                             nN, P(N_row), P(N_ip), P(N_ix), P(N_va), nT, P(T_row), P(T_ip), P(T_ix), P(T_va),
                                             P(bL), bounds_p[0], bounds_p[1], bounds_q[0], bounds_q[1], ctypes.byref(h)))
        self.handle = h
        self.ncol = 12 if params.analytical else 8

    def set_state(self, p=None, q=None, x3=None):
        super().set_state(p=p, q=None if q is None else self.spaces.from_spec(np.asarray(q, dtype=np.float64)), x3=x3)

    def get_state(self):
        p, q, x = super().get_state()
        return p, self.spaces.to_spec(q), x
