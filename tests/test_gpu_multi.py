"""Multi-GPU parity (needs >= 2 CUDA devices; run with `gpurun --gpus 2`): the mesh-partitioned run with
peer-memory halo exchange must reproduce the single-domain CPU oracle on the owned dofs of every rank."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pytestmark = pytest.mark.gpu
K_WAVE, RHO = 3.0, 2.0


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _worker(rank, world, port, scheme, q, flags=0, big=False):
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from oracle.wave_2D_oracle import wave_2D_solve_oracle
        from porthamiltonian_fem_b200 import _lib, partition as pt
        from porthamiltonian_fem_b200.mesh import rectangle_delaunay
        from porthamiltonian_fem_b200.solver import Mesh, Solver, tag_boundary_edges, setup_peer_exchange
        from porthamiltonian_fem_b200.mesh import rectangle_structured
        # big: > 296 patches per rank, so that every CTA of the persistent kernels walks through several patches (stage pipeline,
        # ghost lists arriving one patch ahead) with the exchange in the loop
        mesh = rectangle_structured(200, 50) if big else rectangle_delaunay(48, 12, seed=9)
        steps, tF = (20, 0.004) if big else (40, 0.02)
        ic = scheme.endswith('+IC')          # 'SM+IC': lumped electromechanical model coupled at the left port (wave_2D.py:427-446)
        scheme = scheme.split('+')[0]
        Ho, _, so = wave_2D_solve_oracle(tF, steps, mesh, 1.0, 0.25, timeIntScheme=scheme, K_wave=K_WAVE, rho=RHO, return_state=True,
                                         interConnection='IC' if ic else None)
        part = pt.build_local_part(mesh[0], mesh[1], pt.slab_cell_ranks(mesh[0], mesh[1], world), rank)
        m = Mesh(part.vertices, part.cells, patch_cells=32 if big else 96, vertex_external=part.vertex_external,
                 edge_external=part.edge_external, cell_owned=part.cell_owned)
        m.set_boundary(tag_boundary_edges(m, 'R', 1.0, 0.25))
        prm = _lib.PhwParams()
        prm.scheme, prm.dt, prm.K_wave, prm.rho = _lib.SCHEMES[scheme], tF / steps, K_WAVE, RHO
        prm.input_kind, prm.in_amp, prm.in_omega, prm.in_t_stop = _lib.IN_SINE_WINDOW, 5.0 if ic else 10.0, 8 * np.pi, 0.25
        prm.interconnection = 1 if ic else 0
        if ic:
            from porthamiltonian_fem_b200.wave_2D import _DEFAULT_LUMPED
            for k, v in _DEFAULT_LUMPED.items():
                setattr(prm, k, float(v))
        prm.tol, prm.max_iter, prm.extrap_order, prm.flags = 1e-13, 400, 2, int(flags)
        s = Solver(m, prm, device=rank)
        info = setup_peer_exchange(s, part, dist, rank, world)
        H = s.run(steps)
        paths = s.stats()['kernel_paths']
        if flags & 768:
            assert paths & 12 == 12, 'the pipelined solves did not run: kernel_paths = %d' % paths
        if flags & 2048:
            assert paths & 16, 'the pipelined H_wave pass did not run: kernel_paths = %d' % paths
        if flags & 4096:
            assert paths & 32, 'the pipelined right-hand sides did not run: kernel_paths = %d' % paths
        p, qv, x3 = s.get_state()
        if ic:
            assert np.abs(x3 - so['x']).max() < 1e-10 * max(1.0, np.abs(so['x']).max()), (x3, so['x'])
        ov, oe = part.vertex_owner == rank, part.edge_owner == rank
        topo_edges = so['topo'].edges
        # oracle edge ids of my local edges (through global vertex ids)
        ep = np.linalg.norm(p[ov] - so['p'][part.vertex_gid[ov]]) / np.linalg.norm(so['p'])
        eq = np.linalg.norm(qv[oe] - so['q'][part.edge_gid[oe]]) / np.linalg.norm(so['q'])
        # external copies must equal the owners' values too (halo pushed after the last iteration + local axpy)
        epx = np.abs(p - so['p'][part.vertex_gid]).max() / np.abs(so['p']).max()
        eH = np.abs(H - Ho).max() / np.abs(Ho).max()
        q.put((rank, ep, eq, epx, eH, info))
    finally:
        dist.destroy_process_group()


# flags: 0 lean kernels; 768 pipelined solves (bits 8, 9); 768 + 2048 + 4096 all pipelined kernels (what bench.py runs at N > 1)
@pytest.mark.parametrize('scheme,flags,big', [('SE', 0, False), ('SV', 0, False), ('SE', 768, False), ('SV', 768, False),
                                              ('SE', 768 + 2048 + 4096, False), ('SV', 768 + 2048 + 4096, False),
                                              ('EH', 768 + 2048 + 4096, False),
                                              ('SE', 0, True), ('SE', 768, True), ('SV', 768 + 2048 + 4096, True)])
def test_two_gpu_partitioned_run_matches_oracle(scheme, flags, big):
    _run_partitioned(2, scheme, flags, big)


# implicit schemes (coupled block solve with the exchange inside, skew bound from the distributed power iteration) and the lumped
# model (port fluxes all-reduced; SM: bordered column solved across the ranks) on a partitioned mesh
@pytest.mark.parametrize('scheme', ['SM', 'IE', 'SM+IC', 'SV+IC', 'SE+IC'])
def test_two_gpu_partitioned_implicit_and_lumped_model(scheme):
    _run_partitioned(2, scheme, 0, False)


# four ranks: the interior slabs have two neighbours and two halo columns (patches of halo cells only at the end of the layout)
@pytest.mark.parametrize('scheme,flags,big', [('SE', 768 + 2048 + 4096, True), ('SV', 768 + 2048 + 4096, False), ('SE', 0, False), ('SM+IC', 0, False)])
def test_four_gpu_partitioned_run_matches_oracle(scheme, flags, big):
    _run_partitioned(4, scheme, flags, big)


def _run_partitioned(world, scheme, flags, big):
    if _ngpu() < world:
        pytest.skip('needs %d GPUs' % world)
    import torch.multiprocessing as mp
    sck = socket.socket()
    sck.bind(('127.0.0.1', 0))
    port = sck.getsockname()[1]
    sck.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, scheme, q, flags, big)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ep, eq, epx, eH, info in res:
        assert ep < 1e-10 and eq < 1e-10 and epx < 1e-10 and eH < 1e-10, (rank, ep, eq, epx, eH, info)
