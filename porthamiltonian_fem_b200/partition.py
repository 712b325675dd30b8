"""Mesh partitioning for multi-GPU runs (one process per GPU).

Replaces dolfin's implicit MPI mesh distribution (reference run_main_wave_2D_mpi.sh:1, wave_2D.py:54-61).
Cell-based partition; a vertex / edge dof is owned by the lowest rank among the cells touching it.  The local
mesh of a rank = its cells + every other cell that touches one of its owned dofs (one halo layer), so that all
owned rows can be evaluated locally; dofs of the local mesh owned by another rank are *external*: their values
are pushed by the owner after every operation that changes them (csrc: PeerComm, peer-memory stores over NVLink).

Everything here is host-side integer logic on numpy arrays and is tested on CPU (tests/test_partition_cpu.py,
including a world_size-2 gloo run that emulates the exchange).
"""
import numpy as np


def edge_table(cells, n_vertices):
    """Unique (lo,hi) pairs in lexicographic order - the same canonical edge numbering as libphwave."""
    cells = np.asarray(cells, dtype=np.int64)
    a = cells[:, [1, 2, 0]]
    b = cells[:, [2, 0, 1]]
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    key = lo * n_vertices + hi
    uniq, inv = np.unique(key.ravel(), return_inverse=True)
    edges = np.stack([uniq // n_vertices, uniq % n_vertices], axis=1)
    return edges, inv.reshape(-1, 3)


def slab_cell_ranks(vertices, cells, world, xmin=None, xmax=None):
    """x-slab partition by cell centroid (a 4:1 rectangle cut into x-slabs gives near-square parts)."""
    cx = vertices[np.asarray(cells, dtype=np.int64)][:, :, 0].mean(axis=1)
    xmin = vertices[:, 0].min() if xmin is None else xmin
    xmax = vertices[:, 0].max() if xmax is None else xmax
    r = np.floor((cx - xmin) / (xmax - xmin) * world).astype(np.int64)
    return np.clip(r, 0, world - 1)


def rcb_cell_ranks(vertices, cells, world):
    """Recursive coordinate bisection of the cell centroids into `world` parts (any world >= 1, not only powers of two):
    the partition for general domains (the reference relies on dolfin's graph partitioner, which - like
    metis.h - is not available here).  A part with n ranks is cut along its longer bounding-box axis at the weighted median
    that gives floor(n/2) : ceil(n/2) of its cells to the two halves; ties are broken by cell index, so the result is
    deterministic.  Returns the rank of every cell."""
    cells = np.asarray(cells, dtype=np.int64)
    cen = np.asarray(vertices, dtype=np.float64)[cells].mean(axis=1)
    rank = np.zeros(cells.shape[0], dtype=np.int64)

    def split(idx, r0, n):
        if n <= 1 or idx.size == 0:
            rank[idx] = r0
            return
        lo = cen[idx].min(axis=0)
        hi = cen[idx].max(axis=0)
        axis = 0 if (hi[0] - lo[0]) >= (hi[1] - lo[1]) else 1
        order = idx[np.lexsort((idx, cen[idx, axis]))]
        n_left = n // 2
        k = (idx.size * n_left) // n
        split(order[:k], r0, n_left)
        split(order[k:], r0 + n_left, n - n_left)

    split(np.arange(cells.shape[0], dtype=np.int64), 0, int(world))
    return rank


class LocalPart:
    """What one rank needs: local mesh, global ids, ownership flags, exchange lists (local user numbering)."""
    pass


def build_local_part(vertices, cells, cell_rank, rank):
    """Extract the local mesh of `rank` from a global mesh.

    Returns a LocalPart with
      vertices [nv,2], cells [nc,3] (local vertex ids), cell_gid, vertex_gid, edge_gkey (global (lo,hi) vertex ids),
      vertex_owner, edge_owner (rank owning each local dof), cell_owned (bool),
      send[nb] / recv[nb]: dicts {'v': local vertex ids, 'e': local edge ids} sorted by global id / key, so that
      send lists of one rank and recv lists of its neighbour enumerate the same dofs in the same order."""
    vertices = np.asarray(vertices, dtype=np.float64)
    cells = np.asarray(cells, dtype=np.int64)
    cell_rank = np.asarray(cell_rank, dtype=np.int64)
    NV = vertices.shape[0]
    g_edges, g_cell_edges = edge_table(cells, NV)
    NE = g_edges.shape[0]
    big = np.iinfo(np.int64).max
    v_owner = np.full(NV, big)
    e_owner = np.full(NE, big)
    np.minimum.at(v_owner, cells.ravel(), np.repeat(cell_rank, 3))
    np.minimum.at(e_owner, g_cell_edges.ravel(), np.repeat(cell_rank, 3))
    mine = cell_rank == rank
    touches_owned = (v_owner[cells] == rank).any(axis=1) | (e_owner[g_cell_edges] == rank).any(axis=1)
    keep = mine | touches_owned
    cell_gid = np.nonzero(keep)[0]
    lc = cells[cell_gid]
    vertex_gid = np.unique(lc)
    g2l = -np.ones(NV, dtype=np.int64)
    g2l[vertex_gid] = np.arange(vertex_gid.shape[0])
    part = LocalPart()
    part.rank = rank
    part.vertices = vertices[vertex_gid]
    part.cells = g2l[lc].astype(np.int32)
    part.cell_gid = cell_gid
    part.cell_owned = mine[cell_gid]
    part.vertex_gid = vertex_gid
    part.vertex_owner = v_owner[vertex_gid]
    l_edges, l_cell_edges = edge_table(part.cells, vertex_gid.shape[0])
    # global edge id of every local edge (through global vertex ids; local vertex order preserves global order)
    glo = vertex_gid[l_edges[:, 0]]
    ghi = vertex_gid[l_edges[:, 1]]
    gkey = glo * NV + ghi
    gall = g_edges[:, 0] * NV + g_edges[:, 1]
    geid = np.searchsorted(gall, gkey)
    assert np.array_equal(gall[geid], gkey)
    part.edges = l_edges
    part.edge_gid = geid
    part.edge_owner = e_owner[geid]
    part.n_global_vertices, part.n_global_edges = NV, NE
    _exchange_lists(part)
    return part


def _exchange_lists(part):
    rank = part.rank
    part.send, part.recv = {}, {}
    # recv: external dofs grouped by owner, sorted by global id
    for kind, owner, gid in (('v', part.vertex_owner, part.vertex_gid), ('e', part.edge_owner, part.edge_gid)):
        ext = np.nonzero(owner != rank)[0]
        for nb in np.unique(owner[ext]):
            sel = ext[owner[ext] == nb]
            sel = sel[np.argsort(gid[sel], kind='stable')]
            part.recv.setdefault(int(nb), {})[kind] = sel.astype(np.int32)
    part.vertex_external = (part.vertex_owner != rank).astype(np.uint8)
    part.edge_external = (part.edge_owner != rank).astype(np.uint8)


def needed_by(part_other_recv_gids, part, kind):
    """Local ids (in `part`) of the dofs whose global ids another rank listed in its recv list (same order)."""
    gid = part.vertex_gid if kind == 'v' else part.edge_gid
    order = np.argsort(gid, kind='stable')
    pos = np.searchsorted(gid[order], part_other_recv_gids)
    loc = order[pos]
    assert np.array_equal(gid[loc], part_other_recv_gids)
    return loc.astype(np.int32)


def structured_slab_part(Nx, Ny, world, rank, xLength_per_rank=1.0, yLength=0.25):
    """Local part of rank `rank` of the weak-scaling bench mesh: `world` slabs of Nx x Ny quads side by side
    (global rectangle world*xLength_per_rank x yLength, alternating diagonals), built directly without ever
    forming the global mesh.  Equivalent to build_local_part(global mesh, slab ranks) (tested on small sizes)."""
    from .mesh import rectangle_structured
    if Nx % 2:
        raise ValueError('Nx must be even so that the diagonal pattern is continuous across slabs')
    extra = 1 if rank < world - 1 else 0
    hx = xLength_per_rank / Nx
    v, c = rectangle_structured(Nx + extra, Ny, xLength_per_rank + extra * hx, yLength, x0=rank * xLength_per_rank)
    # exact slab abscissae
    col = np.repeat(np.arange(Nx + extra + 1, dtype=np.int64), Ny + 1)
    row = np.tile(np.arange(Ny + 1, dtype=np.int64), Nx + extra + 1)
    v[:, 0] = rank * xLength_per_rank + col * hx
    gcol = rank * Nx + col
    NVg = (world * Nx + 1) * (Ny + 1)
    part = LocalPart()
    part.rank = rank
    part.vertices, part.cells = v, c
    part.vertex_gid = gcol * (Ny + 1) + row
    ccol = np.repeat(np.arange(Nx + extra, dtype=np.int64), 2 * Ny)
    part.cell_owned = ccol < Nx
    part.cell_gid = None
    # ownership by geometry: the interface line x = rank*L belongs to rank-1; everything right of x = (rank+1)*L
    # belongs to rank+1
    vo = np.full(v.shape[0], rank, dtype=np.int64)
    if rank > 0:
        vo[col == 0] = rank - 1
    vo[col > Nx] = rank + 1
    part.vertex_owner = vo
    edges, _ = edge_table(c, v.shape[0])
    ca, cb = col[edges[:, 0]], col[edges[:, 1]]
    eo = np.full(edges.shape[0], rank, dtype=np.int64)
    if rank > 0:
        eo[(ca == 0) & (cb == 0)] = rank - 1
    eo[np.maximum(ca, cb) > Nx] = rank + 1
    part.edges = edges
    part.edge_owner = eo
    ga, gb = part.vertex_gid[edges[:, 0]], part.vertex_gid[edges[:, 1]]
    part.edge_gid = np.minimum(ga, gb) * NVg + np.maximum(ga, gb)      # global key (not a dense id; only its order matters)
    part.n_global_vertices = NVg
    part.n_global_edges = None
    _exchange_lists(part)
    return part
