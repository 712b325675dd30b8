"""Host-side multi-GPU logic on CPU: partition, ownership, exchange lists; plus a world_size-2 gloo run that
emulates the halo exchange and checks 'partitioned operator == global operator' with the oracle."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import fem
from porthamiltonian_fem_b200.mesh import rectangle_structured, rectangle_delaunay
from porthamiltonian_fem_b200 import partition as pt


@pytest.mark.parametrize('world', [2, 3, 4])
@pytest.mark.parametrize('mesh', ['structured', 'delaunay', 's1c_rcb', 'delaunay_rcb'])
def test_partition_invariants(world, mesh):
    if mesh == 's1c_rcb':
        from porthamiltonian_fem_b200.mesh import square_one_channel
        v, c = square_one_channel(20)
    else:
        v, c = rectangle_structured(24, 6) if mesh == 'structured' else rectangle_delaunay(24, 6, seed=2)
    cr = pt.rcb_cell_ranks(v, c, world) if mesh.endswith('_rcb') else pt.slab_cell_ranks(v, c, world)
    parts = [pt.build_local_part(v, c, cr, r) for r in range(world)]
    NV, NE = parts[0].n_global_vertices, parts[0].n_global_edges
    # every dof is owned by exactly one rank, and every rank holds all cells touching its owned dofs
    own_v = np.zeros(NV, dtype=int)
    own_e = np.zeros(NE, dtype=int)
    for p in parts:
        own_v[p.vertex_gid[p.vertex_owner == p.rank]] += 1
        own_e[p.edge_gid[p.edge_owner == p.rank]] += 1
        assert p.cell_owned.sum() == (cr == p.rank).sum()
    assert np.all(own_v == 1) and np.all(own_e == 1)
    # send/recv symmetry: what r receives from nb is exactly what nb can resolve, in the same order
    for p in parts:
        for nb, lists in p.recv.items():
            for kind in lists:
                gid = (p.vertex_gid if kind == 'v' else p.edge_gid)[lists[kind]]
                loc = pt.needed_by(gid, parts[nb], kind)
                own = (parts[nb].vertex_owner if kind == 'v' else parts[nb].edge_owner)[loc]
                assert np.all(own == nb)
                assert np.all(np.diff(gid) > 0)


@pytest.mark.parametrize('world', [1, 2, 3, 5, 8])
def test_rcb_partition_is_balanced_and_compact(world):
    """recursive coordinate bisection (general domains): balanced to one cell per cut, deterministic, and
    on the unit square with the inlet channel every part is a compact box-like region (few neighbours)"""
    from porthamiltonian_fem_b200.mesh import square_one_channel
    v, c = square_one_channel(20)
    cr = pt.rcb_cell_ranks(v, c, world)
    counts = np.bincount(cr, minlength=world)
    assert counts.sum() == len(c) and counts.min() >= len(c) // world - world and counts.max() <= -(-len(c) // world) + world
    assert np.array_equal(cr, pt.rcb_cell_ranks(v, c, world))
    if world > 1:
        parts = [pt.build_local_part(v, c, cr, r) for r in range(world)]
        assert max(len(p.recv) for p in parts) <= 7            # MAX_RANKS - 1 neighbours fit the device exchange tables
        halo = sum(int((~p.cell_owned).sum()) for p in parts)
        assert halo < 0.6 * len(c)                             # one halo layer around compact parts stays a fraction of the mesh


def test_structured_slab_shortcut_matches_general_partition():
    world, Nx, Ny = 3, 8, 4
    v, c = rectangle_structured(world * Nx, Ny, float(world), 0.25)
    cr = pt.slab_cell_ranks(v, c, world)
    for r in range(world):
        g = pt.build_local_part(v, c, cr, r)
        s = pt.structured_slab_part(Nx, Ny, world, r)
        assert np.array_equal(np.sort(g.vertex_gid), np.sort(s.vertex_gid))
        # same local meshes up to local ordering: compare through global ids
        og, os_ = np.argsort(g.vertex_gid), np.argsort(s.vertex_gid)
        assert np.allclose(g.vertices[og], s.vertices[os_])
        assert np.array_equal(g.vertex_owner[og], s.vertex_owner[os_])
        cg = np.sort(np.sort(g.vertex_gid[g.cells], axis=1), axis=0)
        cs = np.sort(np.sort(s.vertex_gid[s.cells], axis=1), axis=0)
        assert np.array_equal(cg[np.lexsort(cg.T[::-1])], cs[np.lexsort(cs.T[::-1])])
        assert g.cell_owned.sum() == s.cell_owned.sum()
        # edge ownership through global vertex pairs
        def ekeys(p):
            a, b = p.vertex_gid[p.edges[:, 0]], p.vertex_gid[p.edges[:, 1]]
            k = np.minimum(a, b) * 10 ** 9 + np.maximum(a, b)
            o = np.argsort(k)
            return k[o], p.edge_owner[o]
        kg, ogw = ekeys(g)
        ks, osw = ekeys(s)
        assert np.array_equal(kg, ks) and np.array_equal(ogw, osw)
        for nb in g.recv:
            for kind in g.recv[nb]:
                assert len(g.recv[nb][kind]) == len(s.recv[nb][kind])


def _worker(rank, world, port, q, how='slab'):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import torch
        if how == 'rcb_s1c':
            from porthamiltonian_fem_b200.mesh import square_one_channel
            v, c = square_one_channel(20)
            cr = pt.rcb_cell_ranks(v, c, world)
            shape, xL, yL = 'S_1C', 1.0, 0.1
        else:
            v, c = rectangle_delaunay(20, 5, seed=4)
            cr = pt.slab_cell_ranks(v, c, world)
            shape, xL, yL = 'R', 1.0, 0.25
        part = pt.build_local_part(v, c, cr, rank)
        gt = fem.Topology(v, c)
        gops = fem.assemble_wave_operators(gt, fem.tag_boundary(gt, shape, xL, yL))
        rng = np.random.default_rng(0)
        P = rng.standard_normal(gt.n_vertices)
        Q = rng.standard_normal(gt.n_edges)
        lt = fem.Topology(part.vertices, part.cells)
        assert np.array_equal(lt.edges, part.edges)
        # local operators; artificial cut edges carry tag 4 (no boundary term)
        lops = fem.assemble_wave_operators(lt, fem.tag_boundary(lt, shape, xL, yL))
        # local vectors: owned entries from the global vector, external entries filled by the exchange
        p_loc = np.where(part.vertex_owner == rank, P[part.vertex_gid], np.nan)
        q_loc = np.where(part.edge_owner == rank, Q[part.edge_gid], np.nan)
        # exchange: tell neighbours which global ids I need, serve theirs
        need = {nb: {k: (part.vertex_gid if k == 'v' else part.edge_gid)[idx] for k, idx in lists.items()}
                for nb, lists in part.recv.items()}
        allneed = [None] * world
        dist.all_gather_object(allneed, need)
        outbox = {}
        for other in range(world):
            if other == rank or rank not in allneed[other]:
                continue
            outbox[other] = {}
            for kind, gids in allneed[other][rank].items():
                loc = pt.needed_by(gids, part, kind)
                outbox[other][kind] = (p_loc if kind == 'v' else q_loc)[loc]
        allout = [None] * world
        dist.all_gather_object(allout, outbox)
        for nb, lists in part.recv.items():
            for kind, idx in lists.items():
                (p_loc if kind == 'v' else q_loc)[idx] = allout[nb][rank][kind]
        assert not np.isnan(p_loc).any() and not np.isnan(q_loc).any()
        # owned rows of the local operators equal the global rows
        ov, oe = part.vertex_owner == rank, part.edge_owner == rank
        err = 0.0
        err = max(err, np.abs((lops['D'] @ p_loc)[oe] - (gops['D'] @ P)[part.edge_gid[oe]]).max())
        err = max(err, np.abs((lops['D'].T @ q_loc)[ov] - (gops['D'].T @ Q)[part.vertex_gid[ov]]).max())
        err = max(err, np.abs((lops['M_p'] @ p_loc)[ov] - (gops['M_p'] @ P)[part.vertex_gid[ov]]).max())
        err = max(err, np.abs((lops['M_q'] @ q_loc)[oe] - (gops['M_q'] @ Q)[part.edge_gid[oe]]).max())
        # global reductions from owned parts
        e_loc = torch.tensor([float(p_loc[ov] @ (lops['M_p'] @ p_loc)[ov]), float(q_loc[oe] @ (lops['M_q'] @ q_loc)[oe])], dtype=torch.float64)
        dist.all_reduce(e_loc)
        err = max(err, abs(e_loc[0].item() - P @ (gops['M_p'] @ P)) / abs(P @ (gops['M_p'] @ P)))
        err = max(err, abs(e_loc[1].item() - Q @ (gops['M_q'] @ Q)) / abs(Q @ (gops['M_q'] @ Q)))
        q.put((rank, err))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('how,world', [('slab', 2), ('rcb_s1c', 3)])
def test_gloo_world2_partitioned_operators_match_global(how, world):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, how)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err in res:
        assert err < 1e-12, (rank, err)


class _HostOperators:
    """numpy stand-in for the kernel-level entry points a partitioned Solver offers (owned rows of the local operators)"""

    def __init__(self, lops):
        self.lops = lops
        self.lib = self
        self.handle = None
        self._diag = [1.0 / lops['M_p'].diagonal(), 1.0 / lops['M_q'].diagonal()]

    def phw_get_jacobi(self, handle, which, addr):
        import ctypes
        d = self._diag[which]
        ctypes.memmove(addr, np.ascontiguousarray(d).ctypes.data, d.nbytes)
        return 0

    def apply_D(self, p):
        return self.lops['D'] @ p

    def apply_DT(self, q):
        return self.lops['D'].T @ q


def _skew_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from porthamiltonian_fem_b200.solver import distributed_skew_estimate
        v, c = rectangle_delaunay(20, 5, seed=4)
        part = pt.build_local_part(v, c, pt.slab_cell_ranks(v, c, world), rank)
        lt = fem.Topology(part.vertices, part.cells)
        lops = fem.assemble_wave_operators(lt, fem.tag_boundary(lt, 'R', 1.0, 0.25))
        # the Jacobi diagonals of halo dofs are incomplete on this rank, as on the device: the estimate must fetch the owners' values
        need = {nb: {k: (part.vertex_gid if k == 'v' else part.edge_gid)[idx] for k, idx in lists.items()} for nb, lists in part.recv.items()}
        allneed = [None] * world
        dist.all_gather_object(allneed, need)
        send_user = {'v': {}, 'e': {}}
        for other in range(world):
            if other != rank and rank in allneed[other]:
                for kind, gids in allneed[other][rank].items():
                    send_user[kind][other] = pt.needed_by(gids, part, kind)
        sigma = distributed_skew_estimate(_HostOperators(lops), part, dist, rank, world, send_user, iters=300)
        q.put((rank, float(sigma)))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_distributed_skew_estimate_matches_global_operator():
    """implicit schemes in partitioned runs: the power iteration that runs across the ranks (solver.distributed_skew_estimate) finds
    the largest singular value of the GLOBAL Jacobi-scaled skew operator, identically on every rank"""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    import torch.multiprocessing as mp
    v, c = rectangle_delaunay(20, 5, seed=4)
    gt = fem.Topology(v, c)
    g = fem.assemble_wave_operators(gt, fem.tag_boundary(gt, 'R', 1.0, 0.25))
    S = sp.diags(1 / np.sqrt(g['M_q'].diagonal())) @ g['D'] @ sp.diags(1 / np.sqrt(g['M_p'].diagonal()))
    ref = spla.svds(S, k=1, return_singular_vectors=False)[0]
    sck = socket.socket()
    sck.bind(('127.0.0.1', 0))
    port = sck.getsockname()[1]
    sck.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_skew_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1]                       # same bits on both ranks
    assert abs(res[0][1] - ref) < 2e-3 * ref, (res, ref)


def test_halo_only_patches_come_last_and_do_not_size_the_stages():
    """A thin interior slab of an 8-rank strong-scaling run (bench.py, BASELINE configs[3]): its two halo columns form patches of
    neighbour cells only.  They own no row, so the layout puts them behind the patches the kernels process and leaves them out of
    the shared-memory carve-up (every dof of such a patch is a ghost: ~3 x the cells)."""
    from porthamiltonian_fem_b200.solver import Mesh
    part = pt.structured_slab_part(32, 256, 8, 3, 1.0 / 8, 0.25)
    pc = 128
    m = Mesh(part.vertices, part.cells, patch_cells=pc, vertex_external=part.vertex_external, edge_external=part.edge_external,
             cell_owned=part.cell_owned)
    m.self_check()
    cnt = m.patch_counts()                       # own cells, edge-op cells, vertex-op cells, owned vertices, owned edges, ghosts
    owns = (cnt[:, 3] + cnt[:, 4]) > 0
    n_rows = int(owns.sum())
    assert n_rows < len(cnt), 'the slab should have patches made of halo cells only'
    assert owns[:n_rows].all() and not owns[n_rows:].any()
    assert cnt[n_rows:, 5].min() > 2 * pc        # what would have sized the stages
    assert cnt[:n_rows, 5].max() < 2 * pc
    # the stage sizes come from the processed patches: the same as for a layout of this slab's own cells alone, up to its halo
    used, inv = np.unique(part.cells[part.cell_owned.astype(bool)], return_inverse=True)
    own_only = Mesh(part.vertices[used], inv.reshape(-1, 3).astype(np.int32), patch_cells=pc)
    a, b = np.array(m.pipe_smem()), np.array(own_only.pipe_smem())
    assert np.all(a[:5] < 1.35 * b[:5]), (a, b)
    m.close(); own_only.close()
