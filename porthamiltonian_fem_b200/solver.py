"""Thin object wrappers over the C ABI handles (include/phwave.h)."""
import ctypes

import numpy as np

from . import _lib

TAG_LEFT, TAG_TOP, TAG_RIGHT, TAG_GENERAL, TAG_NONE = 0, 1, 2, 3, 4


class Mesh:
    """Connectivity + patch layout of a triangle mesh (phw_mesh_*; host only, no GPU needed)."""

    def __init__(self, vertices, cells, patch_cells=2048):
        self.lib = _lib.load()
        self.vertices = np.ascontiguousarray(vertices, dtype=np.float64)
        self.cells = np.ascontiguousarray(cells, dtype=np.int32)
        if self.vertices.ndim != 2 or self.vertices.shape[1] != 2 or self.cells.ndim != 2 or self.cells.shape[1] != 3:
            raise ValueError('vertices must be [Nv,2], cells must be [Nc,3]')
        h = ctypes.c_void_p()
        _lib.check(self.lib.phw_mesh_create(self.vertices.shape[0], self.vertices.ctypes.data_as(_lib.c_f64p),
                                            self.cells.shape[0], self.cells.ctypes.data_as(_lib.c_i32p),
                                            int(patch_cells), ctypes.byref(h)))
        self.handle = h
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

    def numbering(self):
        v = np.empty(self.n_vertices, dtype=np.int32)
        e = np.empty(self.n_edges, dtype=np.int32)
        c = np.empty(self.n_cells, dtype=np.int32)
        _lib.check(self.lib.phw_mesh_get_numbering(self.handle, v.ctypes.data_as(_lib.c_i32p),
                                                   e.ctypes.data_as(_lib.c_i32p), c.ctypes.data_as(_lib.c_i32p)))
        return v, e, c

    def patch_counts(self):
        c = np.empty((self.n_patches, 6), dtype=np.int32)
        _lib.check(self.lib.phw_mesh_get_patch_counts(self.handle, c.ctypes.data_as(_lib.c_i32p)))
        return c

    def close(self):
        if getattr(self, 'handle', None):
            self.lib.phw_mesh_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def tag_boundary_edges(mesh, domainShape, xLength, yLength):
    """Facet markers of the reference (wave_2D.py:226-280): 0 left/port, 1 top (rectangle only), 2 right,
    3 bottom/general, 4 unmarked.  dolfin marks a facet when its vertices and midpoint satisfy ``inside``;
    later markers override earlier ones (left, top, right, general)."""
    ids, _ = mesh.boundary()
    edges = mesh.edges()[0]
    pa = mesh.vertices[edges[ids, 0]]
    pb = mesh.vertices[edges[ids, 1]]
    pm = 0.5 * (pa + pb)
    tol = 1e-12
    xInputStart = 0.0 if domainShape == 'R' else -0.10

    def on(pred):
        return pred(pa) & pred(pb) & pred(pm)

    tags = np.full(ids.shape[0], TAG_NONE, dtype=np.int32)
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

    def run(self, n_steps, out=None):
        if out is None:
            out = np.zeros((n_steps, self.ncol))
        _lib.check(self.lib.phw_run(self.handle, int(n_steps), _lib.ptr(out)))
        return out

    def stats(self):
        st = _lib.PhwStats()
        _lib.check(self.lib.phw_get_stats(self.handle, ctypes.byref(st)))
        return st.as_dict()

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
