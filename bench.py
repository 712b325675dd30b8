#!/usr/bin/env python
"""bench.py - headline benchmark of the hot path (contract: see the task statement / DESIGN.md "Measurement").

Workload (BASELINE.json configs[3], fits one B200): refined synthetic rectangle 1.0 x 0.25, Nx x Ny quads split
in two -> 25.0 M cells, 50.0 M dofs (P1 vertices + RT1 edges), weak form, symplectic Euler, K=3, rho=2,
dt=2e-5, on the exact standing wave of the reference's analytical case (wave_2D.py:375-377,734) with its left datum
(--from-rest: the reference's default input g(t) = 10 sin(8 pi t) from a zero state, wave_2D.py:384).
One "step" = one time step of the loop wave_2D.py:787-982.  metric = dof-updates/s = N_dof * steps / seconds.

  value : device-timed (CUDA events on the solver's stream inside phw_run, state resident in HBM)
  e2e   : same K steps through the host-buffer C-ABI path (phw_set_state from pinned host memory ->
          phw_run -> H_array rows + phw_get_state back to the host), wall clock
  roofline : the dominant Chebyshev solve kernel: algorithmic bytes / its CUDA-event time; `traffic` = DRAM bytes per launch
          from the committed ncu --set full capture of THAT kernel (profiles/ncu_dram_bytes.json)
  cpu_baseline : the CPU oracle (oracle/, scipy SuperLU factor-once) on a down-scaled mesh of the same family, one
          independent oracle rank per host core (the communication-free upper bound of an MPI run)
  N > 1 : weak scaling (one 50 M-dof slab per GPU) is the headline; the same line carries config.strong = the ONE
          50 M-dof mesh cut into N x-slabs (BASELINE configs[3] as stated), measured in the same run
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K_WAVE, RHO = 3.0, 2.0
FULL_NX, FULL_NY = 7072, 1768
DT_FULL = 2e-5
METRIC = 'dof-updates/sec (fp64, device-timed)'


def clocks_sampler(stop, out, dev):
    q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
    while not stop.is_set():
        try:
            r = subprocess.run(['nvidia-smi', '-i', str(dev), '--query-gpu=' + q, '--format=csv,noheader,nounits'],
                               capture_output=True, text=True, timeout=5)
            f = [x.strip() for x in r.stdout.strip().split(',')]
            if len(f) >= 6:
                out.append(f)
        except Exception:
            pass
        stop.wait(0.2)


def summarize_clocks(samples):
    if not samples:
        return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
    sm = sorted(int(s[0]) for s in samples if s[0].isdigit())
    mx = [int(s[1]) for s in samples if s[1].isdigit()]
    reasons = []
    for idx, name in [(2, 'hw_slowdown'), (3, 'hw_thermal_slowdown'), (4, 'sw_thermal_slowdown'), (5, 'sw_power_cap')]:
        if any(s[idx].lower().startswith('active') for s in samples):
            reasons.append(name)
    return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons}


# ---- CPU arms (the only places that execute oracle/) ---------------------------------------------------------------
def cpu_oracle_rate(nx, ny, steps, refactor, scheme='SE'):
    """dof-updates/s of the CPU oracle (weak, wave only) on an nx x ny x 2 mesh of the same family."""
    from oracle.wave_2D_oracle import wave_2D_solve_oracle
    from porthamiltonian_fem_b200.mesh import rectangle_structured
    mesh = rectangle_structured(nx, ny, 1.0, 0.25)
    dt = DT_FULL * FULL_NX / nx                       # same CFL ratio as the full-size run
    timing = {}
    wave_2D_solve_oracle(dt * steps, steps, mesh, 1.0, 0.25, timeIntScheme=scheme, K_wave=K_WAVE, rho=RHO,
                         refactor_every_solve=refactor, timing=timing)
    n_dof = timing['n_dof']
    return n_dof * steps / timing['loop_seconds'], n_dof, timing['loop_seconds']


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_oracle_rate_all_cores(nx, ny, steps, refactor, scheme='SE', max_procs=128):
    """One independent oracle process per host core, each on its own nx x ny x 2 sample mesh (SuperLU is single-threaded, so
    this is how the reference uses a multi-core host: `mpirun -n <cores>`, run_main_wave_2D_mpi.sh:1 - here without any
    communication, i.e. an upper bound).  Aggregate rate = all dof-updates / the slowest process's loop time."""
    n = max(1, min(host_cores(), max_procs))
    env = dict(os.environ, OMP_NUM_THREADS='1', OPENBLAS_NUM_THREADS='1', MKL_NUM_THREADS='1', CUDA_VISIBLE_DEVICES='')
    cmd = [sys.executable, os.path.abspath(__file__), '--cpu-worker', '{},{},{},{},{}'.format(nx, ny, steps, int(refactor), scheme)]
    procs = [subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, env=env) for _ in range(n)]
    res = []
    for p in procs:
        out, _ = p.communicate()
        for ln in out.splitlines():
            if ln.startswith('{'):
                res.append(json.loads(ln))
    if not res:
        raise RuntimeError('no CPU worker finished')
    secs = max(r['loop_seconds'] for r in res)
    return sum(r['n_dof'] for r in res) * steps / secs, res[0]['n_dof'], secs, len(res)


def run_reference_arm(args):
    """--impl reference: the reference's CPU algorithm (oracle port, re-factorising at every solve call the way
    dolfin's solve(A, x, b) does) on a bounded sample of the workload, on all host cores (one oracle rank per core)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    nx, ny = 221, 55
    steps = max(1, args.steps)
    t0 = time.time()
    rate, n_dof, secs, cores = cpu_oracle_rate_all_cores(nx, ny, steps, True, args.scheme)
    line = {
        'metric': METRIC, 'value': rate, 'unit': 'dof-updates/s', 'impl': 'reference', 'n_gpus': args.gpus,
        'steps': steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * secs / steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': '{} weak P1/RT1 wave on a structured rectangle; CPU sample: {} independent oracle ranks (one per host core), each '
                               '{}x{}x2 cells ({} dofs) of the 7072x1768x2 workload at the same CFL ratio; timed with the host clock around the step '
                               'loop (this arm has no device); the bench arm runs the full 50 M-dof mesh per GPU'.format(args.scheme, cores, nx, ny, n_dof)},
        'cpu_baseline': {'value': rate, 'unit': 'dof-updates/s', 'cores': cores, 'kind': 'port',
                         'sample': '{} steps on {} x ({}x{}x2 cells), scipy SuperLU re-factorised at every solve (dolfin solve(A,x,b) '
                                   'behaviour); the real dolfin/PETSc stack is not installable here'.format(steps, cores, nx, ny)},
        'e2e': {'value': rate, 'unit': 'dof-updates/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'wall_s': time.time() - t0,
    }
    print(json.dumps(line))


def run_sweep_bench(args):
    """--workload sweep: BASELINE configs[4] - 1024 independent IC cases (SV, 104x26x2 = 5408 cells each, parameters from
    default_rng(1234)) advanced together on one GPU (each rank of a multi-GPU run takes its own shard of 1024 cases,
    no collectives).  Same JSON contract, metric = aggregate dof-updates/s."""
    import torch
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    from porthamiltonian_fem_b200.mesh import rectangle_structured
    from porthamiltonian_fem_b200.sweep import wave_sweep_solve, sweep_cases
    n_cases = args.cases
    mesh = rectangle_structured(104, 26, 1.0, 0.25)
    cases = sweep_cases(n_cases * world)[rank * n_cases:(rank + 1) * n_cases]
    dt = 1e-3
    t0 = time.time()
    # --pipe p / q / pq with --patch-cells <= 1024: the bulk-copy pipelined mass solves on the grouped mesh (A/B against the lean kernels)
    sflags = (pipe_flags(args.pipe) & (256 | 512)) if (args.sweep_pipe and 0 < args.patch_cells <= 1024) else 0
    H, nc, solver = wave_sweep_solve(dt * args.warmup, args.warmup, 0, 1.0, 0.25, cases, mesh=mesh, device=local_rank, patch_cells=args.patch_cells if args.patch_cells > 0 else None,
                                     extrap_order=args.extrap, return_solver=True, solver_flags=sflags)
    setup_s = time.time() - t0
    torch.cuda.synchronize()
    H = solver.run(args.steps)
    torch.cuda.synchronize()
    st = solver.stats()
    n_dof_case = (104 + 1) * (26 + 1) + (3 * nc + 2 * (104 + 26)) // 2 + 3
    dev_s = st['device_ms'] * 1e-3
    if world > 1:
        tt = torch.tensor([dev_s], dtype=torch.float64, device='cuda')
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_s = float(tt.item())
    if rank == 0:
        print(json.dumps({'metric': METRIC, 'value': world * n_cases * n_dof_case * args.steps / dev_s,
                          'unit': 'dof-updates/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
                          'ms_per_step': 1e3 * dev_s / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                          'dtype': 'f64', 'data': 'synthetic',
                          'config': {'workload': 'batched sweep: {} independent IC cases per GPU, SV, 104x26x2 cells ({} dofs) each, dt=1e-3'.format(n_cases, n_dof_case),
                                     'iters_per_solve_p': st['iters_p'] / max(1, st['solves_p']), 'iters_per_solve_q': st['iters_q'] / max(1, st['solves_q']),
                                     'max_rel_residual': st['max_rel_residual'], 'setup_seconds': setup_s, 'kernel_paths': int(st['kernel_paths']),
                                     'patches': int(getattr(solver.mesh, 'n_patches', 0)),
                                     'layout': 'state [n_dof][n_cases], case index fastest, one copy of the assembled operators (csrc/kernels_sweep.cuh)' if int(st['kernel_paths']) & 4096
                                               else 'mesh replicated per case (grouped device mesh)'},
                          'gpu_launches': int(st['kernel_launches']),
                          'energy_residual_max': float(np.abs(H[:, :, 2]).max())}))
    if world > 1:
        dist.destroy_process_group()


def ncu_dram_bytes():
    """DRAM bytes per cell and Chebyshev iteration of the solve kernels, from the committed `ncu --set full` captures
    (profiles/ncu_dram_bytes.json, written by profiles/ncu_extract.py from the .ncu-rep of the capture)."""
    try:
        return json.load(open(os.path.join(ROOT, 'profiles', 'ncu_dram_bytes.json')))
    except Exception:
        return {}


def main():
    """Runs the benchmark.  A failure of the pipelined configuration is an error: there is no silent fall-back to other kernels
    (round 1 had one; the default path has since run 12 000 + 20 000 full-size steps without a fault)."""
    return run_bench()


def pipe_flags(pipe):
    return (256 if 'q' in pipe else 0) | (512 if 'p' in pipe else 0) | (2048 if 'e' in pipe else 0) | (4096 if 'r' in pipe else 0)


def build_case(args, torch, dist, rank, world, local_rank, nx_local, x_per_rank):
    """setup (not timed): slab `rank` of a global rectangle (world * x_per_rank) x 0.25, nx_local x ny quads per slab; layout, device
    SoA, peer exchange, state = exact standing wave at phase args.phase (or rest + the reference's default input)."""
    from porthamiltonian_fem_b200 import _lib, partition as pt
    from porthamiltonian_fem_b200.solver import Mesh, Solver, tag_boundary_edges, port_weights, setup_peer_exchange
    t_setup = time.time()
    part = pt.structured_slab_part(nx_local, args.ny, world, rank, x_per_rank, 0.25)
    vertices = part.vertices
    def make_mesh(patch_cells):
        if world > 1:
            return Mesh(part.vertices, part.cells, patch_cells=patch_cells, vertex_external=part.vertex_external,
                        edge_external=part.edge_external, cell_owned=part.cell_owned)
        return Mesh(part.vertices, part.cells, patch_cells=patch_cells)
    xL, yL = x_per_rank * world, 0.25
    omega = np.sqrt(K_WAVE / RHO) * np.sqrt(np.pi ** 2 / (4 * xL ** 2) + 4 * np.pi ** 2 / yL ** 2)

    def tag(m):
        tags = tag_boundary_edges(m, 'R', xL, yL)
        if args.from_rest:
            m.set_boundary(tags)
        else:
            m.set_boundary(tags, port_weights(m, tags, lambda x, y: RHO * omega * np.sin(2 * np.pi * y / yL + np.pi / 2)))
    dt = DT_FULL * FULL_NX * x_per_rank / nx_local     # same CFL ratio as 7072 quads per unit length
    prm = _lib.PhwParams()
    prm.scheme, prm.dt, prm.K_wave, prm.rho = _lib.SCHEMES[args.scheme], dt, K_WAVE, RHO
    if args.from_rest:
        prm.input_kind, prm.in_amp, prm.in_omega, prm.in_t_stop = _lib.IN_SINE_WINDOW, 10.0, 8 * np.pi, 0.25    # wave_2D.py:384
    else:
        prm.input_kind, prm.in_amp, prm.in_omega = _lib.IN_COSINE, 1.0, omega
    prm.tol, prm.max_iter, prm.extrap_order = args.tol, 200, args.extrap
    prm.flags = 4 | (32 if args.cell_rows else 0) | (64 if args.ell_staged else 0) | pipe_flags(args.pipe)
    prm.extrap_order_q = args.extrap_q
    # the two shared-memory stages of the pipelined kernels are sized by the largest patch of the rank's layout: if they do not
    # fit on ANY rank, all ranks step down to the next smaller patch size together (the ranks must agree before the collective setup)
    m = solver = None
    patch_used = None
    for pc in [args.patch_cells] + [c for c in (896, 768, 640, 512) if c < args.patch_cells]:
        ok, err = 1, None
        try:
            m = make_mesh(pc)
            tag(m)
            solver = Solver(m, prm, device=local_rank)
        except _lib.PhwError as ex:
            ok, err = 0, ex
            if 'shared memory' not in str(ex):
                raise
        if world > 1:
            tt = torch.tensor([ok], dtype=torch.int64, device='cuda')
            dist.all_reduce(tt, op=dist.ReduceOp.MIN)
            all_ok = int(tt.item())
        else:
            all_ok = ok
        if all_ok:
            patch_used = pc
            break
        if solver is not None:
            solver.close()
        if m is not None:
            m.close()
        m = solver = None
    if solver is None:
        raise SystemExit('no patch size fits the pipelined kernels on every rank: %r' % (err,))
    n_dof_local = int((part.vertex_owner == rank).sum() + (part.edge_owner == rank).sum())
    n_cells_local = int(part.cell_owned.sum())
    if world > 1:
        setup_peer_exchange(solver, part, dist, rank, world)
        tt = torch.tensor([n_dof_local, n_cells_local], dtype=torch.int64, device='cuda')
        dist.all_reduce(tt)
        n_dof, n_cells_total = int(tt[0].item()), int(tt[1].item())
    else:
        n_dof, n_cells_total = n_dof_local, n_cells_local
    # exact standing wave p = P(x,y) cos(w t), q = -(1/(rho w)) grad P sin(w t) of the analytical case at phase w t0: p nodal, q as RT1
    # edge fluxes (3-point Gauss on every edge).  At w t0 = 0 this is the reference's initial state (q = 0); the default pi/4 is
    # a generic instant of the run, where both fields and both velocities are O(1) of their amplitudes.
    phase = np.pi * args.phase
    kx, ky = np.pi / (2 * xL), 2 * np.pi / yL
    pshape = RHO * omega * np.cos(kx * vertices[:, 0]) * np.cos(ky * vertices[:, 1])      # spatial factor of the exact p at the vertices
    if args.from_rest:
        solver.set_state(p=np.zeros(m.n_vertices), q=np.zeros(m.n_edges), x3=np.zeros(3))
        t0_sim = 0.0
    else:
        q0 = np.zeros(m.n_edges)
        if args.phase != 0.0:
            edges = m.edges()[0]
            a, b = vertices[edges[:, 0]], vertices[edges[:, 1]]
            tx, ty = b[:, 0] - a[:, 0], b[:, 1] - a[:, 1]
            for gp, gw in zip((0.5 - np.sqrt(0.15), 0.5, 0.5 + np.sqrt(0.15)), (5.0 / 18, 8.0 / 18, 5.0 / 18)):
                x, y = a[:, 0] + gp * tx, a[:, 1] + gp * ty
                qx = kx * np.sin(kx * x) * np.cos(ky * y)
                qy = ky * np.cos(kx * x) * np.sin(ky * y)
                q0 += gw * (qx * ty - qy * tx)              # q . (|e| n_e), n_e = tangent (lo -> hi) rotated by -90 degrees
            q0 *= np.sin(phase)
            del edges, a, b, tx, ty, x, y, qx, qy
        solver.set_state(p=pshape * np.cos(phase), q=q0, x3=np.zeros(3))
        t0_sim = phase / omega
        solver.set_time(t0_sim)
        del q0
    return dict(solver=solver, m=m, part=part, n_dof_local=n_dof_local, n_cells_local=n_cells_local, n_dof=n_dof,
                n_cells_total=n_cells_total, omega=omega, dt=dt, t0=t0_sim, pshape=pshape, nx_local=nx_local, x_per_rank=x_per_rank,
                setup_s=time.time() - t_setup, steps_done=0, patch_cells=patch_used)


def run_bench():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='mesh50m', choices=['mesh50m', 'sweep', 'paper'])
    ap.add_argument('--cases', type=int, default=1024)
    ap.add_argument('--sweep-pipe', action='store_true', help='--workload sweep: pipelined mass solves (needs --patch-cells <= 1024)')
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--nx', type=int, default=FULL_NX)
    ap.add_argument('--ny', type=int, default=FULL_NY)
    ap.add_argument('--scheme', default='SE')
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'], help="N > 1: 'weak' (default) one nx x ny slab per GPU as the headline and the "
                    "strong-scaling run of ONE nx x ny mesh in config.strong; 'strong': the strong-scaling run is the headline (value, ms_per_step)")
    ap.add_argument('--no-strong', action='store_true', help='N > 1: skip the strong-scaling part')
    ap.add_argument('--from-rest', action='store_true', help="zero initial state and the reference's default input g(t) = 10 sin(8 pi t), t < 0.25 (wave_2D.py:384)")
    ap.add_argument('--patch-cells', type=int, default=0, help='cells per patch (0: 1024 with the pipelined M_q solve so that two stages fit in shared memory, else 2048)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--tol', type=float, default=1e-13)
    ap.add_argument('--extrap', type=int, default=3)
    ap.add_argument('--extrap-q', type=int, default=0, help='extrapolation order of the q solves (0: same as --extrap)')
    ap.add_argument('--phase', type=float, default=0.25, help='phase omega*t0 / pi of the standing wave the run starts from (0: the t = 0 state of wave_2D.py:734)')
    ap.add_argument('--ell-staged', action='store_true', help='ELL rows of the M_p solve staged in shared memory by cp.async (A/B)')
    ap.add_argument('--cell-rows', action='store_true', help='cell-based instead of assembled (ELL) rows in the M_p solve (A/B)')
    ap.add_argument('--pipe', default='auto', help="'auto' (default) = 'pqer': all bulk-copy pipelined kernels (csrc/kernels_pipe.cuh), on one GPU and in partitioned runs; "
                    "'none': lean kernels; or any of 'p' (M_p solve), 'q' (M_q solve), 'e' (fused H_wave reduction), 'r' (right-hand sides)")
    ap.add_argument('--stream-probe', action='store_true', help='print the bandwidth of a pure streaming pass over the ELL M_p data (diagnostics)')
    ap.add_argument('--debug-residuals', action='store_true', help='print the residual history of the last M_p / M_q solve to stderr')
    ap.add_argument('--cpu-worker', default=None, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.cpu_worker:      # one oracle rank of the CPU arms (see cpu_oracle_rate_all_cores)
        nx, ny, steps, refactor, scheme = args.cpu_worker.split(',')
        rate, n_dof, secs = cpu_oracle_rate(int(nx), int(ny), int(steps), bool(int(refactor)), scheme)
        print(json.dumps({'rate': rate, 'n_dof': n_dof, 'loop_seconds': secs}))
        return
    if args.pipe == 'auto':
        args.pipe = 'pqer'
    if args.pipe == 'none':
        args.pipe = ''
    if args.patch_cells <= 0 and args.workload == 'mesh50m':
        args.patch_cells = 1024 if 'q' in args.pipe else 2048
    if args.impl == 'reference':
        return run_reference_arm(args)
    if args.workload == 'sweep':
        return run_sweep_bench(args)
    if args.workload == 'paper':
        from profiles.paper_cases import run_paper_bench
        return run_paper_bench(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (no CPU fallback)')
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    from porthamiltonian_fem_b200 import _lib

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if world == 1:
            return x
        tt = torch.tensor([x], dtype=torch.float64, device='cuda')
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    def allsum(xs):
        if world == 1:
            return list(xs)
        tt = torch.tensor(list(xs), dtype=torch.float64, device='cuda')
        dist.all_reduce(tt)
        return [float(v) for v in tt.tolist()]

    def measure(case, steps, warmup, sample_clocks):
        """W warm-up steps, then EXACTLY `steps` steps bracketed by barrier + synchronize; device time = CUDA events inside
        phw_run on the solver's stream, max over ranks"""
        solver = case['solver']
        barrier()
        solver.run(warmup)
        barrier()
        samples, stop = [], threading.Event()
        th = None
        if sample_clocks:
            th = threading.Thread(target=clocks_sampler, args=(stop, samples, local_rank), daemon=True)
            th.start()
        barrier()
        t0 = time.perf_counter()
        solver.run(steps)
        barrier()
        wall = time.perf_counter() - t0
        stop.set()
        if th:
            th.join(timeout=2)
        st = solver.stats()
        case['steps_done'] += warmup + steps
        return dict(dev_s=allmax(st['device_ms'] * 1e-3), st=st, samples=samples, wall=wall)

    # ---- weak configuration (headline): every rank owns one slab of nx x ny quads of a (world x 1.0) x 0.25 rectangle
    weak = build_case(args, torch, dist, rank, world, local_rank, args.nx, 1.0)
    solver, m = weak['solver'], weak['m']
    r = measure(weak, args.steps, args.warmup, True)
    st, dev_s, samples, wall = r['st'], r['dev_s'], r['samples'], r['wall']
    clocks = summarize_clocks(samples)
    if world > 1:      # every rank samples its own GPU: a power-capped straggler shows up here (all ranks advance in lock-step)
        allclk = [None] * world
        dist.all_gather_object(allclk, clocks)
        clocks = dict(clocks, per_rank=allclk, sm_mhz_min_over_ranks=min((c['sm_mhz'] or 0) for c in allclk),
                      reasons=sorted(set(x for c in allclk for x in c['reasons'])))
    if args.stream_probe and rank == 0:
        for nt, occ in ((256, 8), (256, 4), (512, 4), (1024, 2), (128, 16), (256, 2)):
            ms, nb = solver.stream_probe(nt, occ, 10)
            sys.stderr.write('stream probe {} threads x {} CTAs/SM: {:.4f} ms per pass, {:.3f} GB, {:.0f} GB/s\n'.format(nt, occ, ms, nb / 1e9, nb / ms / 1e6))
    if world > 1 and (int(os.environ.get('PHW_XCH_DEBUG', '0')) & 64):
        # exchange timing of the last M_p / M_q solve (diagnostics): per iteration, in us, how long CTA 0 waited at the grid barrier,
        # how late the last CTA arrived after CTA 0, the local reduction, and the wait for the other ranks' norms
        ts = np.zeros((2, 32, 5), dtype=np.uint64)
        _lib.check(_lib.load().phw_debug_xch_times(solver.handle, ts.ctypes.data))
        for f, nm in ((0, 'M_p'), (1, 'M_q')):
            t = ts[f].astype(np.float64)
            ok = t[:, 3] > 0
            t = t[ok][:12]
            if len(t) > 2:
                d = lambda a, b: ' '.join('%5.1f' % (1e-3 * x) for x in (t[:, a] - t[:, b]))
                sys.stderr.write('rank %d %s us barrier-wait(CTA0): %s | last arrival after CTA0: %s | reduce: %s | peer wait: %s | iteration: %s\n' % (
                    rank, nm, d(1, 0), d(4, 0), d(2, 1), d(3, 2), ' '.join('%5.1f' % (1e-3 * x) for x in np.diff(t[:, 0]))))
    if args.debug_residuals and rank == 0:
        for w, nm in ((0, 'M_p'), (1, 'M_q')):
            sys.stderr.write('residual history {}: {}\n'.format(nm, ' '.join('%.1e' % r_ for r_ in solver.residual_history(w))))
    n_dof = weak['n_dof']
    value = n_dof * args.steps / dev_s

    # ---- end-to-end through host buffers (pinned), copies inside the timed region
    p_host = torch.zeros(m.n_vertices, dtype=torch.float64).pin_memory()
    q_host = torch.zeros(m.n_edges, dtype=torch.float64).pin_memory()
    x_host = np.zeros(3)
    H_host = torch.zeros((args.steps, 8), dtype=torch.float64).pin_memory()
    p_np, q_np = p_host.numpy(), q_host.numpy()
    lib = _lib.load()
    _lib.check(lib.phw_get_state(solver.handle, p_np.ctypes.data, q_np.ctypes.data, x_host.ctypes.data))
    barrier()
    t0 = time.perf_counter()
    _lib.check(lib.phw_set_state(solver.handle, p_np.ctypes.data, q_np.ctypes.data, x_host.ctypes.data))
    _lib.check(lib.phw_run(solver.handle, args.steps, H_host.numpy().ctypes.data))
    _lib.check(lib.phw_get_state(solver.handle, p_np.ctypes.data, q_np.ctypes.data, x_host.ctypes.data))
    barrier()
    e2e_s = allmax(time.perf_counter() - t0)
    weak['steps_done'] += args.steps
    e2e_value = n_dof * args.steps / e2e_s
    nloc = m.n_vertices + m.n_edges
    h2d = world * (nloc * 8 + 24) / args.steps
    d2h = world * (nloc * 8 + 24 + args.steps * 8 * 8) / args.steps
    # ---- correctness at bench size: the state just read back against the exact standing wave at the simulated time (nodal values
    # of p on this rank's vertices; first-order scheme, so the expected level is O(omega dt) for SE and O((omega dt)^2) for SV)
    exact = None
    if not args.from_rest:
        t_end = weak['t0'] + weak['steps_done'] * weak['dt']
        pe = weak['pshape'] * np.cos(weak['omega'] * t_end)
        d = p_np - pe
        s2, e2, finite = allsum([float(d @ d), float(pe @ pe), float(np.isfinite(p_np).all() and np.isfinite(q_np).all())])
        emax, pmax = allmax(float(np.abs(d).max())), allmax(float(np.abs(pe).max()))
        exact = {'p_nodal_rel_l2': float(np.sqrt(s2 / e2)), 'p_nodal_max_over_amplitude': emax / pmax, 'all_finite': finite == world,
                 'steps_from_exact_start': weak['steps_done'], 'omega_dt': weak['omega'] * weak['dt'],
                 'note': 'exact p = rho w cos(pi x / 2 xL) cos(2 pi y / yL) cos(w t) (wave_2D.py:734-741); the scheme is O(w dt) (SE) / O((w dt)^2) (SV) accurate'}
        del pe, d
    del p_host, q_host, p_np, q_np

    # ---- strong scaling of ONE nx x ny mesh over the N GPUs (BASELINE configs[3] as stated; role of run_main_wave_2D_mpi.sh:1)
    strong = None
    if world > 1 and not args.no_strong and args.nx % (2 * world) == 0:
        sc = build_case(args, torch, dist, rank, world, local_rank, args.nx // world, 1.0 / world)
        rs = measure(sc, args.steps, max(args.warmup, 5), False)
        ms1, ms1_src = None, None
        try:     # the 1-GPU time of the same mesh: the last committed 1-GPU line of this bench
            ref = json.load(open(os.path.join(ROOT, 'profiles', 'bench_1gpu_reference.json')))
            if ref.get('config', {}).get('dofs_total') == sc['n_dof'] and ref.get('steps'):
                ms1, ms1_src = float(ref['ms_per_step']), 'profiles/bench_1gpu_reference.json (committed 1-GPU line of this bench, same mesh)'
        except Exception:
            pass
        ms_strong = 1e3 * rs['dev_s'] / args.steps
        ms_weak = 1e3 * dev_s / args.steps
        strong = {'dofs_total': sc['n_dof'], 'dofs_per_gpu': sc['n_dof_local'], 'ms_per_step': ms_strong, 'value': sc['n_dof'] * args.steps / rs['dev_s'],
                  'iters_per_solve_p': rs['st']['iters_p'] / max(1, rs['st']['solves_p']), 'iters_per_solve_q': rs['st']['iters_q'] / max(1, rs['st']['solves_q']),
                  'ms_per_iter_p': rs['st']['ms_solve_p'] / max(1, rs['st']['iters_p']), 'ms_per_iter_q': rs['st']['ms_solve_q'] / max(1, rs['st']['iters_q']),
                  'max_rel_residual': rs['st']['max_rel_residual'], 'setup_seconds': sc['setup_s'], 'patch_cells': sc['patch_cells'],
                  'efficiency_vs_weak_slab': ms_weak / (world * ms_strong),
                  'efficiency_vs_weak_slab_note': 'T(one 50 M-dof slab per GPU, this run, exchange included) / (N x T(50 M dofs over N GPUs))',
                  'ms_per_step_1gpu': ms1, 'ms_per_step_1gpu_source': ms1_src, 'efficiency': (ms1 / (world * ms_strong)) if ms1 else None}
        sc['solver'].close()
        sc['m'].close()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        peak = float(peaks.get('hbm_gbs', 6650.0))
        peak_src = 'measured (MEASURED_PEAKS.json hbm_gbs)' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s (B200_PROFILING.md)'
        # the two solve kernels take ~85 % of the step; the one with the larger share of the timed region is reported as
        # the dominant kernel.  Algorithmic bytes per cell and Chebyshev iteration: 40 (M_p), 96 (M_q).
        nc = weak['n_cells_local']
        paths = int(st.get('kernel_paths', 0))      # phw_path bits: 1 lean M_p, 2 lean M_q, 4 pipelined M_p, 8 pipelined M_q
        ncu = ncu_dram_bytes()
        kern = {}
        for key, short, alg, it, ms, solves in (
                ('p', 'solve_ell_pipe_k' if paths & 4 else 'solve_ell_coop_k' if paths & 1 else 'solve_coop_k<0>', 40.0, st['iters_p'], st['ms_solve_p'], st['solves_p']),
                ('q', 'solve_mq_pipe_k' if paths & 8 else 'solve_mq_coop_k' if paths & 2 else 'solve_coop_k<1>', 96.0, st['iters_q'], st['ms_solve_q'], st['solves_q'])):
            ach = alg * nc * it / (ms * 1e-3) / 1e9 if ms > 0 else None
            nb = ncu.get(short, {})
            kern[key] = {'kernel': short + ' (Chebyshev M_{} solve, all iterations in one cooperative launch)'.format(key), 'achieved': ach,
                         'frac': (ach / peak) if ach else None, 'share_of_step': ms / st['device_ms'],
                         'bytes_per_launch': alg * nc * it / max(1, solves), 'ms_per_launch': ms / max(1, solves),
                         'ms_per_iteration': ms / max(1, it),
                         'traffic': (nb['dram_bytes_per_cell_iteration'] * nc * it / max(1, solves)) if nb else None,
                         'traffic_source': nb.get('source') if nb else None}
        dom = kern['p'] if kern['p']['share_of_step'] >= kern['q']['share_of_step'] else kern['q']
        oth = kern['q'] if dom is kern['p'] else kern['p']
        line = {
            'metric': METRIC, 'value': value, 'unit': 'dof-updates/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dev_s / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': '{} weak P1/RT1 wave, {}, structured rectangle {}x{}x2 = {} cells and {} dofs per GPU ({} dofs in total, x-slab partition, peer-memory halo exchange '
                                   'fused into the solve kernels), dt={:g}; state vectors ({:.0f} MB each) exceed L2, no flush needed'.format(
                                       args.scheme,
                                       "zero state + the reference's default input 10 sin(8 pi t) (wave_2D.py:384)" if args.from_rest else
                                       'exact standing wave of the analytical case at phase w*t0 = {:.3g} pi with its left datum (wave_2D.py:375-377,734)'.format(args.phase),
                                       args.nx, args.ny, nc, weak['n_dof_local'], n_dof, weak['dt'], 8e-6 * m.n_edges),
                       'dofs_per_gpu': weak['n_dof_local'], 'dofs_total': n_dof, 'patch_cells': weak['patch_cells'], 'phase_over_pi': args.phase, 'pipe': args.pipe, 'tol': args.tol, 'extrap_order': args.extrap,
                       'iters_per_solve_p': st['iters_p'] / max(1, st['solves_p']), 'iters_per_solve_q': st['iters_q'] / max(1, st['solves_q']),
                       'max_rel_residual': st['max_rel_residual'], 'setup_seconds': weak['setup_s'], 'exact_error': exact, 'strong': strong},
            'e2e': {'value': e2e_value, 'unit': 'dof-updates/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'seconds': e2e_s},
            'gpu_launches': int(st['kernel_launches']),
            'clocks': clocks,
            'roofline': {'bound': 'hbm', 'kernel': dom['kernel'], 'achieved': dom['achieved'], 'peak': peak, 'unit': 'GB/s',
                         'frac': dom['frac'], 'traffic': dom['traffic'], 'peak_source': peak_src, 'traffic_source': dom['traffic_source'],
                         'bytes_per_launch': dom['bytes_per_launch'], 'ms_per_launch': dom['ms_per_launch'],
                         'kernel_share_of_step': dom['share_of_step'], 'second_kernel': oth,
                         'whole_loop_model_gbs': st['bytes_model'] / (st['device_ms'] * 1e-3) / 1e9, 'whole_loop_model_frac': st['bytes_model'] / (st['device_ms'] * 1e-3) / 1e9 / peak,
                         'ms_per_iter_q': kern['q']['ms_per_iteration'], 'ms_per_iter_p': kern['p']['ms_per_iteration']},
            'wall_s_timed_region': wall,
        }
        if args.scaling == 'strong' and strong:
            line.update({'value': strong['value'], 'ms_per_step': strong['ms_per_step'], 'scaling': 'strong'})
            line['config']['weak'] = {'value': value, 'ms_per_step': 1e3 * dev_s / args.steps}
            line['config']['dofs_total'], line['config']['dofs_per_gpu'] = strong['dofs_total'], strong['dofs_per_gpu']
        if not args.no_cpu_baseline:
            try:
                t_cpu = time.time()
                rate, nd, secs, cores = cpu_oracle_rate_all_cores(442, 110, 40, False, args.scheme)
                line['cpu_baseline'] = {'value': rate, 'unit': 'dof-updates/s', 'cores': cores, 'kind': 'port',
                                        'sample': '40 {} steps on {} x (442x110x2-cell mesh, {} dofs) of the same family: one CPU-oracle rank per host core '
                                                  '(scipy SuperLU is single-threaded; factor once: {:.1f} s in the slowest step loop, {:.1f} s with assembly and '
                                                  'factorisation); one core alone: {:.3g} dof-updates/s'.format(args.scheme, cores, nd, secs, time.time() - t_cpu, rate / cores)}
            except Exception as ex:   # the bench line must still be printed
                line['cpu_baseline'] = {'value': None, 'unit': 'dof-updates/s', 'cores': host_cores(), 'kind': 'port', 'sample': 'failed: %r' % (ex,)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
