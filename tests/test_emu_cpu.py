"""CPU check of the bulk-copy pipelined kernels (csrc/kernels_pipe.cuh) inside a host-side SPMD emulator
(tests/emu: one OS thread per CTA, one fiber per CUDA thread, asynchronous copies deferred until they are waited for, guard pages
behind every array and behind the dynamic shared memory).  The SAME kernel sources are compiled with g++; only the handful of
inline-PTX helpers is replaced (PHW_HOST_EMU).  What this covers without a GPU: indexing and shared-memory carve-ups, the staging
protocol (every stage is waited for before it is read, expected byte counts match, 16-byte rules of the bulk copies, nothing in
flight at exit), the Chebyshev control flow and the arithmetic - against the oracle's assembled operators and sparse direct
solves.  What it does not cover: real asynchrony, memory-model effects, performance (the `-m gpu` suite and bench.py do that)."""
import ctypes
import os
import sys

import numpy as np
import pytest
import scipy.sparse.linalg as spla

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'emu'))

from oracle import fem
from porthamiltonian_fem_b200.mesh import rectangle_structured, rectangle_delaunay, square_one_channel
from porthamiltonian_fem_b200.solver import Mesh, tag_boundary_edges


@pytest.fixture(scope='module')
def emu():
    from build_emu import build
    return ctypes.CDLL(build())


def run(lib, what, v, c, pc, grid, in1, in2=None, nout=None, tags=None, tol=1e-13):
    v = np.ascontiguousarray(v, dtype=np.float64)
    c = np.ascontiguousarray(c, dtype=np.int32)
    out = np.zeros(nout)
    it = ctypes.c_int32(0)
    err = ctypes.create_string_buffer(512)
    in1 = np.ascontiguousarray(in1, dtype=np.float64)
    p2 = pt = None
    if in2 is not None:
        in2 = np.ascontiguousarray(in2, dtype=np.float64)
        p2 = in2.ctypes.data_as(ctypes.c_void_p)
    if tags is not None:
        tags = np.ascontiguousarray(tags, dtype=np.int32)
        pt = tags.ctypes.data_as(ctypes.c_void_p)
    rc = lib.emu_run(what.encode(), ctypes.c_int64(len(v)), v.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(len(c)),
                     c.ctypes.data_as(ctypes.c_void_p), ctypes.c_int32(pc), ctypes.c_int32(grid), pt,
                     in1.ctypes.data_as(ctypes.c_void_p), p2, out.ctypes.data_as(ctypes.c_void_p), ctypes.byref(it),
                     ctypes.c_double(tol), err, 512)
    if rc:
        raise RuntimeError(err.value.decode())
    return out, it.value


def rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


MESHES = [
    # name, mesh, shape, xL, yL, patch_cells, CTAs: several patches per CTA, odd / even starts of the owned ranges, one patch, two cells
    ('delaunay_60x15_pc48_g4', lambda: rectangle_delaunay(60, 15, seed=5), 'R', 1.0, 0.25, 48, 4),
    ('structured_40x10_pc64_g3', lambda: rectangle_structured(40, 10), 'R', 1.0, 0.25, 64, 3),
    ('s1c_20_pc40_g5', lambda: square_one_channel(20), 'S_1C', 1.0, 0.1, 40, 5),
    ('single_patch', lambda: rectangle_structured(10, 4), 'R', 1.0, 0.25, 4096, 2),
    ('tiny_2x1', lambda: rectangle_structured(2, 1), 'R', 1.0, 0.25, 16, 2),
]


@pytest.mark.parametrize('name,make,shape,xL,yL,pc,grid', MESHES)
def test_emulated_mass_solves_match_direct_solves(emu, name, make, shape, xL, yL, pc, grid):
    v, c = make()
    topo = fem.Topology(v, c)
    ops = fem.assemble_wave_operators(topo, fem.tag_boundary(topo, shape, xL, yL))
    rng = np.random.default_rng(2)
    bq = rng.standard_normal(topo.n_edges)
    bp = rng.standard_normal(topo.n_vertices)
    xq = spla.spsolve(ops['M_q'].tocsc(), bq)
    xp = spla.spsolve(ops['M_p'].tocsc(), bp)
    its = {}
    for what, b, ref in [('mq_lean', bq, xq), ('mq1', bq, xq), ('mp_lean', bp, xp), ('mp1', bp, xp)]:
        x, its[what] = run(emu, what, v, c, pc, grid, b, nout=len(b))
        assert rel(x, ref) < 1e-11, (what, rel(x, ref))
    # the lean and the pipelined kernels run the same iteration (the stop test may differ by one step)
    assert abs(its['mq1'] - its['mq_lean']) <= 1 and abs(its['mp1'] - its['mp_lean']) <= 1


@pytest.mark.parametrize('name,make,shape,xL,yL,pc,grid', MESHES)
def test_emulated_energy_and_right_hand_sides_match_oracle(emu, name, make, shape, xL, yL, pc, grid):
    v, c = make()
    topo = fem.Topology(v, c)
    ops = fem.assemble_wave_operators(topo, fem.tag_boundary(topo, shape, xL, yL))
    m = Mesh(v, c, patch_cells=pc)
    tags = tag_boundary_edges(m, shape, xL, yL)
    m.close()
    rng = np.random.default_rng(3)
    p = rng.standard_normal(topo.n_vertices)
    q = rng.standard_normal(topo.n_edges)
    en, _ = run(emu, 'energy', v, c, pc, grid, p, q, nout=2)
    assert abs(en[0] - p @ (ops['M_p'] @ p)) < 1e-13 * abs(en[0]) and abs(en[1] - q @ (ops['M_q'] @ q)) < 1e-13 * abs(en[1])
    y, _ = run(emu, 'rhs_e', v, c, pc, grid, p, nout=topo.n_edges, tags=tags)
    assert rel(y, ops['D'] @ p) < 1e-14
    y, _ = run(emu, 'rhs_v', v, c, pc, grid, q, nout=topo.n_vertices, tags=tags)
    assert rel(y, ops['D'].T @ q) < 1e-14


def test_emulated_solves_with_zero_right_hand_side_and_loose_tolerance(emu):
    v, c = rectangle_delaunay(30, 8, seed=7)
    topo = fem.Topology(v, c)
    for what, n in [('mq1', topo.n_edges), ('mp1', topo.n_vertices)]:
        x, it = run(emu, what, v, c, 40, 3, np.zeros(n), nout=n)
        assert np.all(x == 0.0) and it <= 2
    ops = fem.assemble_wave_operators(topo, fem.tag_boundary(topo, 'R', 1.0, 0.25))
    b = np.random.default_rng(4).standard_normal(topo.n_edges)
    x6, it6 = run(emu, 'mq1', v, c, 40, 3, b, nout=len(b), tol=1e-6)
    x13, it13 = run(emu, 'mq1', v, c, 40, 3, b, nout=len(b), tol=1e-13)
    assert it6 < it13 and rel(x6, spla.spsolve(ops['M_q'].tocsc(), b)) < 1e-5


@pytest.mark.parametrize('what', ['mq1', 'mp1', 'energy', 'rhs_e'])
def test_a_stage_that_never_arrives_ends_in_a_clean_error(what):
    """failure path of the pipelined kernels (pipe_stage_ready): a lost bulk copy must end in an error code - all CTAs leave the
    kernel (no hang at the grid barrier), nobody computes on the stage that did not arrive.  Runs in a child process because the
    fault injection of the emulator is switched on by an environment variable read once."""
    import subprocess, textwrap
    code = textwrap.dedent('''
        import sys, ctypes, numpy as np
        sys.path.insert(0, {tests!r}); sys.path.insert(0, {emu!r}); sys.path.insert(0, {root!r})
        from test_emu_cpu import run
        from build_emu import build
        from porthamiltonian_fem_b200.mesh import rectangle_delaunay
        lib = ctypes.CDLL(build())
        v, c = rectangle_delaunay(40, 10, seed=1)
        big = np.ones(10 ** 5)                      # longer than any dof vector of this mesh: the harness reads what it needs
        try:
            run(lib, {what!r}, v, c, 40, 3, big, big if {what!r} == 'energy' else None, nout=10 ** 5)
            print('NO ERROR')
        except RuntimeError as ex:
            print('ERROR:', ex)
    ''').format(tests=HERE, emu=os.path.join(HERE, 'emu'), root=os.path.dirname(HERE), what=what)
    env = dict(os.environ, PHW_EMU_DROP='7')
    r = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert 'ERROR: a stage did not arrive' in r.stdout, r.stdout + r.stderr


def _run2(lib, what, v, c, pc, grid, b_global, kind, tol=1e-13):
    """two ranks in one process: x-slab partition, peer exchange through each other's buffers and mailboxes (emu_run2)"""
    from porthamiltonian_fem_b200 import partition as pt
    cr = pt.slab_cell_ranks(v, c, 2)
    parts = [pt.build_local_part(v, c, cr, r) for r in range(2)]
    keep = []

    def A(x, dt):
        x = np.ascontiguousarray(x, dtype=dt)
        keep.append(x)
        return x

    def ptrs(arrs):
        return (ctypes.c_void_p * len(arrs))(*[a.ctypes.data_as(ctypes.c_void_p) for a in arrs])
    vtx = [A(p.vertices, np.float64) for p in parts]
    cells = [A(p.cells, np.int32) for p in parts]
    vext = [A(p.vertex_external, np.uint8) for p in parts]
    eext = [A(p.edge_external, np.uint8) for p in parts]
    gid = [(p.edge_gid if kind == 'e' else p.vertex_gid) for p in parts]
    b = [A(b_global[g], np.float64) for g in gid]
    out = [A(np.zeros(len(g)), np.float64) for g in gid]
    sloc, srem = [], []
    for r in range(2):      # rank r pushes what the other rank lists in its recv[r]: my local ids / the other's local ids, same order
        lst = parts[1 - r].recv.get(r, {}).get(kind, np.zeros(0, np.int32))
        g = (parts[1 - r].edge_gid if kind == 'e' else parts[1 - r].vertex_gid)[lst]
        sloc.append(A(pt.needed_by(g, parts[r], kind), np.int32))
        srem.append(A(lst, np.int32))
    nv = (ctypes.c_int64 * 2)(*[len(p.vertices) for p in parts])
    nc = (ctypes.c_int64 * 2)(*[len(p.cells) for p in parts])
    ne = (ctypes.c_int64 * 2)(*[len(p.edge_gid) for p in parts])
    ns = (ctypes.c_int32 * 2)(*[len(x) for x in sloc])
    it = (ctypes.c_int32 * 2)()
    err = ctypes.create_string_buffer(512)
    rc = lib.emu_run2(what.encode(), ctypes.c_int32(pc), ctypes.c_int32(grid), nv, ptrs(vtx), nc, ptrs(cells), ptrs(vext), ne, ptrs(eext),
                      ptrs(b), ptrs(out), ns, ptrs(sloc), ptrs(srem), it, ctypes.c_double(tol), err, 512, None, None, None)
    if rc:
        raise RuntimeError(err.value.decode())
    x = np.full(len(b_global), np.nan)
    for r, p in enumerate(parts):
        own = (p.edge_owner if kind == 'e' else p.vertex_owner) == r
        x[gid[r][own]] = out[r][own]
    ext_off = max(np.abs(out[r] - x[gid[r]]).max() for r in range(2))
    return x, list(it), ext_off


@pytest.mark.parametrize('mesh', ['delaunay', 'structured'])
def test_emulated_two_rank_solves_with_peer_exchange(emu, mesh):
    """partitioned run of the pipelined solves with the inter-GPU exchange (PHW_PIPE_MULTI; not yet run on 2 GPUs) next
    to the lean kernels that carry the same exchange on the GPU: owned parts reproduce the global solve, both ranks take the same
    stop decision, and the external copies equal their owners' values"""
    v, c = rectangle_delaunay(40, 10, seed=3) if mesh == 'delaunay' else rectangle_structured(36, 9)
    topo = fem.Topology(v, c)
    ops = fem.assemble_wave_operators(topo, fem.tag_boundary(topo, 'R', 1.0, 0.25))
    rng = np.random.default_rng(1)
    bq = rng.standard_normal(topo.n_edges)
    bp = rng.standard_normal(topo.n_vertices)
    xq = spla.spsolve(ops['M_q'].tocsc(), bq)
    xp = spla.spsolve(ops['M_p'].tocsc(), bp)
    its = {}
    for what, b, ref, kind in [('mq_lean', bq, xq, 'e'), ('mq1', bq, xq, 'e'), ('mp_lean', bp, xp, 'v'), ('mp1', bp, xp, 'v')]:
        x, it, ext_off = _run2(emu, what, v, c, 40, 3, b, kind)
        assert not np.isnan(x).any() and rel(x, ref) < 1e-11, (what, rel(x, ref))
        assert it[0] == it[1] and ext_off == 0.0, (what, it, ext_off)
        its[what] = it[0]
    assert abs(its['mq1'] - its['mq_lean']) <= 1 and abs(its['mp1'] - its['mp_lean']) <= 1


def _run2_single(lib, what, v, c, pc, grid, p_global, q_global, shape='R', xL=1.0, yL=0.25):
    """energy / right-hand-side kernels on the two partitioned layouts of an x-slab split (no exchange inside these kernels: the
    external entries of the inputs are simply given their owners' values, as the solve's exchange and the local state updates
    leave them).  Returns per-rank outputs and the parts."""
    from porthamiltonian_fem_b200 import partition as pt
    cr = pt.slab_cell_ranks(v, c, 2)
    parts = [pt.build_local_part(v, c, cr, r) for r in range(2)]
    keep = []

    def A(x, dt):
        x = np.ascontiguousarray(x, dtype=dt)
        keep.append(x)
        return x

    def ptrs(arrs):
        return (ctypes.c_void_p * len(arrs))(*[a.ctypes.data_as(ctypes.c_void_p) for a in arrs])
    vtx = [A(p.vertices, np.float64) for p in parts]
    cells = [A(p.cells, np.int32) for p in parts]
    vext = [A(p.vertex_external, np.uint8) for p in parts]
    eext = [A(p.edge_external, np.uint8) for p in parts]
    owned = [A(p.cell_owned, np.uint8) for p in parts]
    tags = []
    for p in parts:
        m = Mesh(p.vertices, p.cells, patch_cells=pc, vertex_external=p.vertex_external, edge_external=p.edge_external)
        tags.append(A(tag_boundary_edges(m, shape, xL, yL), np.int32))
        m.close()
    if what == 'energy':
        in1 = [A(p_global[p.vertex_gid], np.float64) for p in parts]
        in2 = [A(q_global[p.edge_gid], np.float64) for p in parts]
        out = [A(np.zeros(2), np.float64) for _ in parts]
    elif what == 'rhs_e':
        in1 = [A(p_global[p.vertex_gid], np.float64) for p in parts]
        in2 = in1
        out = [A(np.full(len(p.edge_gid), np.nan), np.float64) for p in parts]
    else:
        in1 = [A(q_global[p.edge_gid], np.float64) for p in parts]
        in2 = in1
        out = [A(np.full(len(p.vertex_gid), np.nan), np.float64) for p in parts]
    empty = [A(np.zeros(1, np.int32), np.int32) for _ in parts]
    nv = (ctypes.c_int64 * 2)(*[len(p.vertices) for p in parts])
    nc = (ctypes.c_int64 * 2)(*[len(p.cells) for p in parts])
    ne = (ctypes.c_int64 * 2)(*[len(p.edge_gid) for p in parts])
    ns = (ctypes.c_int32 * 2)(0, 0)
    it = (ctypes.c_int32 * 2)()
    err = ctypes.create_string_buffer(512)
    rc = lib.emu_run2(what.encode(), ctypes.c_int32(pc), ctypes.c_int32(grid), nv, ptrs(vtx), nc, ptrs(cells), ptrs(vext), ne, ptrs(eext),
                      ptrs(in1), ptrs(out), ns, ptrs(empty), ptrs(empty), it, ctypes.c_double(1e-13), err, 512,
                      ptrs(owned), ptrs(tags), ptrs(in2))
    if rc:
        raise RuntimeError(err.value.decode())
    return out, parts


@pytest.mark.parametrize('mesh', ['delaunay', 'structured'])
def test_emulated_two_rank_energy_and_right_hand_sides(emu, mesh):
    """the pipelined H_wave pass and right-hand sides on partitioned layouts (round 2: they now also run in multi-GPU runs): the two
    ranks' energy shares add up to the global quadratic forms (every cell counted once: halo cells of the other rank carry the
    'foreign' flag), and every owned row equals the global operator's row"""
    v, c = rectangle_delaunay(40, 10, seed=3) if mesh == 'delaunay' else rectangle_structured(36, 9)
    topo = fem.Topology(v, c)
    ops = fem.assemble_wave_operators(topo, fem.tag_boundary(topo, 'R', 1.0, 0.25))
    rng = np.random.default_rng(5)
    p = rng.standard_normal(topo.n_vertices)
    q = rng.standard_normal(topo.n_edges)
    out, parts = _run2_single(emu, 'energy', v, c, 40, 3, p, q)
    ep, eq = out[0][0] + out[1][0], out[0][1] + out[1][1]
    assert abs(ep - p @ (ops['M_p'] @ p)) < 1e-13 * abs(ep) and abs(eq - q @ (ops['M_q'] @ q)) < 1e-13 * abs(eq)
    assert min(out[0][0], out[1][0]) > 0.2 * ep          # both ranks hold a real share
    for what, ref, kind in [('rhs_e', ops['D'] @ p, 'e'), ('rhs_v', ops['D'].T @ q, 'v')]:
        out, parts = _run2_single(emu, what, v, c, 40, 3, p, q)
        for r, part in enumerate(parts):
            own = (part.edge_owner if kind == 'e' else part.vertex_owner) == r
            gid = part.edge_gid if kind == 'e' else part.vertex_gid
            assert own.sum() > 0 and rel(out[r][own], ref[gid[own]]) < 1e-14, (what, r)


def _run_csr(lib, what, M, B, cpl, mode, n_groups, x, cscale=None, alpha=1.0, lam=(1.0, 1.0), tol=1e-13):
    M = M.tocsr()
    M.sort_indices()
    ip = np.ascontiguousarray(M.indptr, dtype=np.int32)
    ix = np.ascontiguousarray(M.indices, dtype=np.int32)
    va = np.ascontiguousarray(M.data, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.zeros(B if what == 'dot' else (M.shape[0], B))
    it = ctypes.c_int32(0)
    err = ctypes.create_string_buffer(512)
    cs = None if cscale is None else np.ascontiguousarray(cscale, dtype=np.float64)
    rc = lib.emu_run_csr(what.encode(), ctypes.c_int32(M.shape[0]), ctypes.c_int32(M.shape[1]), ip.ctypes.data_as(ctypes.c_void_p),
                         ix.ctypes.data_as(ctypes.c_void_p), va.ctypes.data_as(ctypes.c_void_p), ctypes.c_int32(B), ctypes.c_int32(cpl),
                         ctypes.c_int32(mode), ctypes.c_int32(n_groups), x.ctypes.data_as(ctypes.c_void_p),
                         None if cs is None else cs.ctypes.data_as(ctypes.c_void_p), ctypes.c_double(alpha), ctypes.c_double(lam[0]),
                         ctypes.c_double(lam[1]), ctypes.c_double(tol), out.ctypes.data_as(ctypes.c_void_p), ctypes.byref(it), err, 512)
    if rc:
        raise RuntimeError(err.value.decode())
    return out, it.value


@pytest.mark.parametrize('B,cpl', [(7, 1), (10, 2), (12, 4), (40, 4), (66, 2)])
@pytest.mark.parametrize('mode', [0, 1])
def test_emulated_sweep_kernels_case_index_fastest(emu, B, cpl, mode):
    """csrc/kernels_sweep.cuh on the CPU: ONE assembled operator, B cases with the case index fastest, 1 / 2 / 4 cases per lane (16-byte
    accesses), both row mappings, chunks that are not full (guard pages behind every array): right-hand sides, per-case quadratic
    forms and the persistent Chebyshev solve against numpy / sparse LU per case"""
    import scipy.sparse.linalg as spla
    v, c = rectangle_structured(6, 2)
    topo = fem.Topology(v, c)
    ops = fem.assemble_wave_operators(topo, fem.tag_boundary(topo, 'R', 1.0, 0.25))
    rng = np.random.default_rng(3)
    Mq, D = ops['M_q'].tocsr(), ops['D'].tocsr()
    ne, nv = D.shape
    n_groups = 2
    # right-hand side of the q rows: (1 / rho_c) D p
    P = rng.standard_normal((nv, B))
    cs = rng.uniform(0.5, 2.0, B)
    Y, _ = _run_csr(emu, 'store', D, B, cpl, mode, n_groups, P, cscale=cs, alpha=-1.5)
    assert np.abs(Y - (-1.5) * (D @ P) * cs[None, :]).max() < 1e-13 * np.abs(D @ P).max()
    # per-case energies 0.5 K_c q^T M_q q
    Q = rng.standard_normal((ne, B))
    E, _ = _run_csr(emu, 'dot', Mq, B, cpl, mode, n_groups, Q, cscale=cs, alpha=0.5)
    assert np.abs(E - 0.5 * cs * np.einsum('ic,ic->c', Q, Mq @ Q)).max() < 1e-12 * np.abs(E).max()
    # all Chebyshev iterations of M_q X = Bm for the B cases in one launch
    d = Mq.diagonal()
    S = (Mq.toarray() / np.sqrt(d)[:, None]) / np.sqrt(d)[None, :]
    ev = np.linalg.eigvalsh(S)
    Bm = rng.standard_normal((ne, B))
    X, its = _run_csr(emu, 'solve', Mq, B, cpl, mode, n_groups, Bm, lam=(0.98 * ev[0], 1.02 * ev[-1]))
    ref = spla.splu(Mq.tocsc()).solve(Bm)
    assert np.abs(X - ref).max() < 1e-11 * np.abs(ref).max()
    assert 5 < its < 80


def _ellipse_for_rectangle(alpha, beta, mu):
    """Python copy of ellipse_for_rectangle (csrc/phwave.cu): Chebyshev ellipse through the corner of [alpha, beta] x i[-mu, mu]"""
    d, delta = 0.5 * (alpha + beta), 0.5 * (beta - alpha)
    if mu <= 0:
        return d, delta
    best, best_c = 1e300, delta
    for i in range(1, 4001):
        a = delta * (1.0 + 4.0 * i / 4000.0 * (mu / delta + 0.05))
        t = 1.0 - delta * delta / (a * a)
        if t <= 0 or a >= d:
            continue
        b = mu / np.sqrt(t)
        c2 = a * a - b * b
        if c2 < 0:
            continue
        rate = (a + b) / (d + np.sqrt(d * d - c2))
        if rate < best:
            best, best_c = rate, np.sqrt(c2)
    return d, best_c


@pytest.mark.parametrize('order,theta', [((1, 1), 0.5), ((2, 2), 0.5), ((3, 2), 1.0)])
def test_emulated_block_solve_of_the_implicit_schemes_on_csr_rows(emu, order, theta):
    """solve_block_csr_k (IE / SM at basis_order > 1, wave_2D.py:305-319,560-572) on the CPU: the coupled system
    [[M_p, th dt K D^T], [-th dt/rho D, M_q]] assembled by the oracle's higher-order spaces, against a sparse direct solve"""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    from oracle import fem_ho
    v, c = rectangle_structured(6, 2)
    topo = fem.Topology(v, c)
    ops = fem_ho.SpacesHO(topo, *order).assemble(fem.tag_boundary(topo, 'R', 1.0, 0.25))
    Mp, Mq, D = ops['M_p'].tocsr(), ops['M_q'].tocsr(), ops['D'].tocsr()
    DT = D.T.tocsr()
    K, rho, dt = 3.0, 2.0, 2e-3
    cDv, cDe = theta * dt * K, -theta * dt / rho
    dp, dq = Mp.diagonal(), Mq.diagonal()
    lam = [np.linalg.eigvalsh((M.toarray() / np.sqrt(d)[:, None]) / np.sqrt(d)[None, :]) for M, d in ((Mp, dp), (Mq, dq))]
    sig = np.linalg.svd((D.toarray() / np.sqrt(dq)[:, None]) / np.sqrt(dp)[None, :], compute_uv=False)[0]
    mu = 1.05 * theta * dt * np.sqrt(K / rho) * sig
    cd, cc = _ellipse_for_rectangle(min(lam[0][0], lam[1][0]), max(lam[0][-1], lam[1][-1]), mu)
    rng = np.random.default_rng(8)
    bp, bq = rng.standard_normal(Mp.shape[0]), rng.standard_normal(Mq.shape[0])
    mats = []
    for M in (Mp, Mq, D, DT):
        M.sort_indices()
        mats.append((np.ascontiguousarray(M.indptr, dtype=np.int32), np.ascontiguousarray(M.indices, dtype=np.int32),
                     np.ascontiguousarray(M.data, dtype=np.float64)))
    P3 = ctypes.c_void_p * 4
    xp, xq = np.zeros(Mp.shape[0]), np.zeros(Mq.shape[0])
    it = ctypes.c_int32(0)
    err = ctypes.create_string_buffer(512)
    rc = emu.emu_run_block(ctypes.c_int32(Mp.shape[0]), ctypes.c_int32(Mq.shape[0]), P3(*[m[0].ctypes.data for m in mats]),
                           P3(*[m[1].ctypes.data for m in mats]), P3(*[m[2].ctypes.data for m in mats]),
                           ctypes.c_double(cDv), ctypes.c_double(cDe), ctypes.c_double(1.0 / K), ctypes.c_double(rho),
                           ctypes.c_double(cd), ctypes.c_double(cc), ctypes.c_double(1e-13), ctypes.c_int32(3),
                           bp.ctypes.data_as(ctypes.c_void_p), bq.ctypes.data_as(ctypes.c_void_p), xp.ctypes.data_as(ctypes.c_void_p),
                           xq.ctypes.data_as(ctypes.c_void_p), ctypes.byref(it), err, 512)
    assert rc == 0, err.value.decode()
    Abig = sp.bmat([[Mp, cDv * DT], [cDe * D, Mq]], format='csc')
    ref = spla.spsolve(Abig, np.concatenate([bp, bq]))
    got = np.concatenate([xp, xq])
    assert np.linalg.norm(got - ref) < 1e-10 * np.linalg.norm(ref), (np.linalg.norm(got - ref) / np.linalg.norm(ref), it.value)
    assert 5 < it.value < 400


def _sliver_mesh():
    """a quality mesh with a band of stretched cells (aspect ratio ~25), as in tests/test_gpu_parity.py"""
    v, c = rectangle_structured(40, 10)
    v = v.copy()
    x = v[:, 0].copy()
    h = 1.0 / 40
    band = (x > 0.5 - 1e-9) & (x < 0.5 + 3 * h + 1e-9)
    v[band, 0] = 0.5 + (x[band] - 0.5) * 0.04
    right = x > 0.5 + 3 * h + 1e-9
    v[right, 0] = 0.5 + 3 * h * 0.04 + (x[right] - (0.5 + 3 * h)) * (0.5 - 3 * h * 0.04) / (0.5 - 3 * h)
    return v, c


@pytest.mark.parametrize('name,make,pc,grid', [('aspect12', lambda: rectangle_structured(6, 18), 64, 3), ('sliver_band', _sliver_mesh, 96, 4),
                                               ('delaunay', lambda: rectangle_delaunay(30, 8, seed=5), 48, 4)])
def test_emulated_conjugate_gradient_mass_solves(emu, name, make, pc, grid):
    """solve_pcg_coop_k (meshes of poor quality) on the CPU: both mass solves against sparse LU; on the mesh with a band of slivers CG
    needs about as many iterations as the host CG of the same matrix (103), far fewer than the element-wise Chebyshev interval implies"""
    import scipy.sparse.linalg as spla
    v, c = make()
    topo = fem.Topology(v, c)
    ops = fem.assemble_wave_operators(topo, fem.tag_boundary(topo, 'R', 1.0, 0.25))
    rng = np.random.default_rng(5)
    bp, bq = rng.standard_normal(topo.n_vertices), rng.standard_normal(topo.n_edges)
    xq, itq = run(emu, 'mq_pcg', v, c, pc, grid, bq, nout=topo.n_edges)
    xp, itp = run(emu, 'mp_pcg', v, c, pc, grid, bp, nout=topo.n_vertices)
    assert rel(xq, spla.spsolve(ops['M_q'].tocsc(), bq)) < 1e-10, (name, itq)
    assert rel(xp, spla.spsolve(ops['M_p'].tocsc(), bp)) < 1e-11, (name, itp)
    assert 5 < itp < 40
    if name == 'sliver_band':
        assert itq < 130, itq
