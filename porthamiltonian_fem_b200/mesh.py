"""Mesh front-end: the stand-in for ``mshr.generate_mesh`` (reference wave_2D.py:93-119).

The reference meshes its domains with mshr/CGAL, which is not available (and not
reproducible) here, so the solver accepts *any* conforming triangle mesh as
``(vertices float64[Nv,2], cells int32[Nc,3])`` and this module supplies generators:

* ``rectangle_structured``  - Nx x Ny quads split by alternating diagonals (the synthetic
  "refined rectangle" family of BASELINE.json, and the mesh the oracle KATs are quoted on);
* ``rectangle_delaunay``    - jittered-lattice Delaunay triangulation (unstructured stand-in);
* ``square_one_channel``    - the 'S_1C' domain of wave_2D.py:97-119 (unit square plus inlet
  channel, locally refined near the inlet);
* ``generate_mesh``         - maps the reference's ``(domainShape, nx, xLength, yLength)`` to
  one of the above with a comparable cell count (mshr gives ~1.46 nx^2 cells on 1 x 0.25);
* ``load_mesh`` / ``save_mesh`` - triangulations from / to files: ``.npz`` (arrays ``vertices``, ``cells``), a pair of
  ``.npy`` files, or the XDMF + raw-binary / inline-XML layout dolfin's ``XDMFFile.write(mesh)`` family uses for
  triangle meshes (``Topology`` ``Triangle`` + ``Geometry`` ``XY``; HDF5 heavy data needs h5py, which this image
  lacks - such files are rejected with a clear message).  Lets a mesh produced by mshr on another machine
  (``wave_2D.py:93-119``) run here unchanged: ``wave_2D_solve(..., mesh=load_mesh(path))``.
"""
import numpy as np


def rectangle_structured(Nx, Ny, xLength=1.0, yLength=0.25, x0=0.0, y0=0.0):
    """Nx x Ny quads on [x0,x0+xL] x [y0,y0+yL]; vertex id = i*(Ny+1)+j; quad (i,j) with corners
    a=(i,j) b=(i+1,j) c=(i+1,j+1) d=(i,j+1) is split (a,b,c),(a,c,d) when i+j is even and
    (a,b,d),(b,c,d) otherwise.  Cell id = 2*(i*Ny+j)+{0,1}."""
    Nx, Ny = int(Nx), int(Ny)
    xs = x0 + xLength * np.arange(Nx + 1, dtype=np.float64) / Nx
    ys = y0 + yLength * np.arange(Ny + 1, dtype=np.float64) / Ny
    vertices = np.empty(((Nx + 1) * (Ny + 1), 2), dtype=np.float64)
    vertices[:, 0] = np.repeat(xs, Ny + 1)
    vertices[:, 1] = np.tile(ys, Nx + 1)
    i = np.repeat(np.arange(Nx, dtype=np.int64), Ny)
    j = np.tile(np.arange(Ny, dtype=np.int64), Nx)
    a = i * (Ny + 1) + j
    b = a + (Ny + 1)
    c = b + 1
    d = a + 1
    even = ((i + j) % 2) == 0
    t0 = np.where(even[:, None], np.stack([a, b, c], 1), np.stack([a, b, d], 1))
    t1 = np.where(even[:, None], np.stack([a, c, d], 1), np.stack([b, c, d], 1))
    cells = np.empty((2 * Nx * Ny, 3), dtype=np.int32)
    cells[0::2] = t0
    cells[1::2] = t1
    return vertices, cells


def _delaunay_cells(points):
    from scipy.spatial import Delaunay
    tri = Delaunay(points)
    cells = tri.simplices.astype(np.int32)
    p = points[cells]
    area2 = (p[:, 1, 0] - p[:, 0, 0]) * (p[:, 2, 1] - p[:, 0, 1]) - (p[:, 1, 1] - p[:, 0, 1]) * (p[:, 2, 0] - p[:, 0, 0])
    scale = np.max(np.abs(area2))
    return cells[np.abs(area2) > 1e-12 * scale]          # drop degenerate slivers on the hull


def rectangle_delaunay(Nx, Ny, xLength=1.0, yLength=0.25, jitter=0.25, seed=0):
    """Delaunay triangulation of a lattice whose interior points are jittered by `jitter`*h."""
    rng = np.random.default_rng(seed)
    X, Y = np.meshgrid(np.arange(Nx + 1, dtype=np.float64), np.arange(Ny + 1, dtype=np.float64), indexing='ij')
    interior_x = (X > 0) & (X < Nx)
    interior_y = (Y > 0) & (Y < Ny)
    X = X + np.where(interior_x, jitter * rng.uniform(-1, 1, X.shape), 0.0)
    Y = Y + np.where(interior_y, jitter * rng.uniform(-1, 1, Y.shape), 0.0)
    pts = np.stack([X.ravel() * (xLength / Nx), Y.ravel() * (yLength / Ny)], axis=1)
    return pts, _delaunay_cells(pts)


def square_one_channel(N, xLength=1.0, yLength=0.1, refine=True):
    """'S_1C' domain (wave_2D.py:97-119): [0,xL]^2 plus the inlet channel
    [-0.1,0] x [xL/2-yL/2, xL/2+yL/2], lattice spacing h = xL/N (N must make 0.05/h integral so the
    channel walls lie on lattice lines), with the lattice halved within distance yL of the inlet
    mouth (0, xL/2) to mimic the reference's single local ``refine`` pass (wave_2D.py:108-119)."""
    h = xLength / N
    nch = 0.1 / h
    nhalf = (yLength / 2) / h
    if abs(nch - round(nch)) > 1e-9 or abs(nhalf - round(nhalf)) > 1e-9:
        raise ValueError('S_1C lattice: N must make 0.1/h and yLength/(2h) integers')
    ylo, yhi = xLength / 2 - yLength / 2, xLength / 2 + yLength / 2

    def inside(x, y, tol=1e-12):
        main = (x >= -tol) & (x <= xLength + tol) & (y >= -tol) & (y <= xLength + tol)
        chan = (x >= -0.1 - tol) & (x <= tol) & (y >= ylo - tol) & (y <= yhi + tol)
        return main | chan

    gx = np.arange(-int(round(nch)), N + 1) * h
    gy = np.arange(0, N + 1) * h
    X, Y = np.meshgrid(gx, gy, indexing='ij')
    pts = np.stack([X.ravel(), Y.ravel()], axis=1)
    pts = pts[inside(pts[:, 0], pts[:, 1])]
    if refine:
        hf = h / 2
        r = yLength + h
        fx = np.arange(-int(round(nch)) * 2, int(np.ceil(r / hf)) + 1) * hf
        fy = xLength / 2 + np.arange(-int(np.ceil(r / hf)), int(np.ceil(r / hf)) + 1) * hf
        FX, FY = np.meshgrid(fx, fy, indexing='ij')
        fp = np.stack([FX.ravel(), FY.ravel()], axis=1)
        keep = inside(fp[:, 0], fp[:, 1]) & (np.hypot(fp[:, 0], fp[:, 1] - xLength / 2) <= yLength + 1e-12)
        fp = fp[keep]
        pts = np.unique(np.round(np.concatenate([pts, fp]) / (hf / 4)).astype(np.int64), axis=0) * (hf / 4)
    # snap to exact lattice values (avoid 1e-17 noise on boundary coordinates)
    pts = np.round(pts / (h / 8)) * (h / 8)
    cells = _delaunay_cells(pts)
    cen = pts[cells].mean(axis=1)
    cells = cells[inside(cen[:, 0], cen[:, 1], tol=-1e-9)]
    # drop unreferenced points
    used = np.unique(cells)
    remap = -np.ones(pts.shape[0], dtype=np.int64)
    remap[used] = np.arange(used.shape[0])
    return pts[used], remap[cells].astype(np.int32)


def structured_dims_for_nx(nx, xLength, yLength):
    """Quad counts giving about the cell count of mshr at resolution nx (~1.46 nx^2 triangles on the
    1 x 0.25 rectangle; mshr's cell size scales with the bounding-box diagonal over nx)."""
    diag = np.hypot(xLength, yLength)
    hq = 0.5677 * diag / nx
    Ny = max(1, int(round(yLength / hq)))
    Nx = max(1, int(round(xLength / hq)))
    return Nx, Ny


def generate_mesh(domainShape, nx, xLength, yLength, kind='structured', seed=0):
    """Stand-in for the mesh block of wave_2D.py:93-123.  Returns (vertices, cells)."""
    if domainShape == 'R':
        Nx, Ny = structured_dims_for_nx(nx, xLength, yLength)
        if kind == 'structured':
            return rectangle_structured(Nx, Ny, xLength, yLength)
        if kind == 'delaunay':
            return rectangle_delaunay(Nx, Ny, xLength, yLength, seed=seed)
        raise ValueError('unknown mesh kind {}'.format(kind))
    if domainShape == 'S_1C':
        N = 20 * max(1, int(round(1.3 * nx / 20.0)))
        return square_one_channel(N, xLength, yLength)
    print('domainShape \"{}\" hasn\'t been created'.format(domainShape))
    raise SystemExit


def _check_mesh(vertices, cells):
    vertices = np.ascontiguousarray(vertices, dtype=np.float64)
    cells = np.ascontiguousarray(cells)
    if vertices.ndim != 2 or vertices.shape[1] not in (2, 3) or cells.ndim != 2 or cells.shape[1] != 3:
        raise ValueError('a mesh is (vertices [Nv,2], cells [Nc,3]); got {} and {}'.format(vertices.shape, cells.shape))
    if vertices.shape[1] == 3:
        if np.abs(vertices[:, 2]).max(initial=0.0) > 0:
            raise ValueError('only planar meshes (z = 0) are supported')
        vertices = np.ascontiguousarray(vertices[:, :2])
    if cells.size and (cells.min() < 0 or cells.max() >= vertices.shape[0]):
        raise ValueError('cell vertex index out of range')
    return vertices, cells.astype(np.int32)


def save_mesh(path, vertices, cells):
    """Write ``path`` (.npz) with the arrays ``vertices`` and ``cells``."""
    vertices, cells = _check_mesh(vertices, cells)
    np.savez(path, vertices=vertices, cells=cells)


def load_mesh(path, cells_path=None):
    """(vertices, cells) from ``.npz`` | (``vertices.npy``, ``cells.npy``) | ``.xdmf`` (Triangle topology, XY geometry with
    inline XML or raw-binary heavy data)."""
    import os
    ext = os.path.splitext(path)[1].lower()
    if ext == '.npz':
        with np.load(path) as z:
            return _check_mesh(z['vertices'], z['cells'])
    if ext == '.npy':
        if cells_path is None:
            raise ValueError('load_mesh(vertices.npy, cells.npy): the cells file is missing')
        return _check_mesh(np.load(path), np.load(cells_path))
    if ext == '.xdmf':
        import xml.etree.ElementTree as ET
        root = ET.parse(path).getroot()
        grid = next((g for g in root.iter('Grid') if g.find('Topology') is not None and g.find('Geometry') is not None), None)
        if grid is None:
            raise ValueError('{}: no Grid with Topology and Geometry'.format(path))
        topo, geom = grid.find('Topology'), grid.find('Geometry')
        if topo.get('TopologyType', topo.get('Type', '')).lower() != 'triangle':
            raise ValueError('{}: only Triangle topologies are supported'.format(path))

        def read_item(item, dtype):
            dims = [int(x) for x in item.get('Dimensions').split()]
            fmt = item.get('Format', 'XML').upper()
            text = (item.text or '').strip()
            if fmt == 'XML':
                return np.array(text.split(), dtype=dtype).reshape(dims)
            if fmt == 'BINARY':
                prec = int(item.get('Precision', '8' if dtype == np.float64 else '4'))
                kind = {('f', 8): np.float64, ('f', 4): np.float32, ('i', 8): np.int64, ('i', 4): np.int32}[('f' if dtype == np.float64 else 'i', prec)]
                raw = np.fromfile(os.path.join(os.path.dirname(os.path.abspath(path)), text), dtype=kind,
                                  count=int(np.prod(dims)), offset=int(item.get('Seek', '0')))
                return raw.reshape(dims).astype(dtype)
            raise ValueError('{}: heavy data in {} format needs h5py, which is not available here; re-export the mesh as .npz '
                             '(numpy.savez(path, vertices=mesh.coordinates(), cells=mesh.cells()))'.format(path, fmt))
        cells = read_item(topo.find('DataItem'), np.int64)
        vertices = read_item(geom.find('DataItem'), np.float64)
        return _check_mesh(vertices, cells)
    raise ValueError('unknown mesh file type {!r} (use .npz, .npy + .npy or .xdmf)'.format(ext))
