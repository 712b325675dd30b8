"""The paper-size configurations (BASELINE.json configs[0..2]; main_wave_2D.py:24-36,92-102,119-120, main_wave_2D_control.py:31-41,
97-98) through the drop-in wave_2D_solve call on one GPU, with the CPU oracle (factor-once) beside it.  These meshes are
cache-resident (<= 44 k dofs): the numbers are latency-, not bandwidth-bound; by default they run the single-launch whole-loop
kernel (csrc/kernels_loop.cuh), with PHW_NO_LOOP=1 the multi-launch path of round 1.

    python profiles/paper_cases.py            one JSON line per case
    python bench.py --workload paper          the same cases as ONE bench line (aggregate dof-updates/s + the per-case list)
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

K, RHO = 3.0, 2.0


def cases():
    from porthamiltonian_fem_b200.mesh import rectangle_structured, generate_mesh
    return [
        ('config1 wave SE nx80-equivalent 136x34x2', rectangle_structured(136, 34), 'R', 1.0, 0.25, 1.5, 3000, dict(timeIntScheme='SE')),
        ('config2 IC SV nx120-equivalent 208x52x2', rectangle_structured(208, 52), 'R', 1.0, 0.25, 4.5, 3000, dict(timeIntScheme='SV', interConnection='IC')),
        ('config2 IC SM nx120-equivalent 208x52x2', rectangle_structured(208, 52), 'R', 1.0, 0.25, 4.5, 3000, dict(timeIntScheme='SM', interConnection='IC')),
        ('config2 IC SV S_1C nx60', generate_mesh('S_1C', 60, 1.0, 0.1), 'S_1C', 1.0, 0.1, 8.0, 16000, dict(timeIntScheme='SV', interConnection='IC')),
        ('config3 passivity SM nx40-equivalent', generate_mesh('R', 40, 1.0, 0.25), 'R', 1.0, 0.25, 8.0, 8000,
         dict(timeIntScheme='SM', interConnection='IC', controlType='passivity', input_stop_t=0.0, control_start_t=1.0)),
        ('config3 casimir SM xLength=4 nx100-equivalent', generate_mesh('R', 100, 4.0, 0.25), 'R', 4.0, 0.25, 5.0, 5000,
         dict(timeIntScheme='SM', controlType='casimir')),
    ]


def run_case(name, mesh, shape, xL, yL, tF, steps, kw, cpu_steps):
    from porthamiltonian_fem_b200 import wave_2D_solve
    t0 = time.time()
    H, nc, st = wave_2D_solve(tF, steps, '/tmp', 0, xL, yL, domainShape=shape, K_wave=K, rho=RHO, mesh=mesh, return_state=True, verbose=False, **kw)
    wall = time.time() - t0
    stats = st['stats']
    ndof = len(st['p']) + len(st['q']) + (3 if kw.get('interConnection') else 0)
    out = {'case': name, 'cells': nc, 'dofs': ndof, 'steps': steps, 'patches': st['mesh'].n_patches, 'gpu_device_ms': stats['device_ms'],
           'gpu_steps_per_s': steps / (stats['device_ms'] * 1e-3), 'gpu_us_per_step': 1e3 * stats['device_ms'] / steps,
           'gpu_dof_updates_per_s': ndof * steps / (stats['device_ms'] * 1e-3),
           'gpu_wall_s_incl_setup': wall, 'launches': stats['kernel_launches'], 'whole_loop_kernel': bool(stats['kernel_paths'] & 256), 'resident_solves': bool(stats['kernel_paths'] & 512),
           'iters_p': stats['iters_p'] / max(1, stats['solves_p']), 'iters_q': stats['iters_q'] / max(1, stats['solves_q']),
           'iters_block': stats['iters_block'] / max(1, stats['solves_block']), 'max_rel_residual': stats['max_rel_residual'],
           'max_abs_E': float(np.abs(H[:, 2]).max()), 'H_end': float(H[-1, 1])}
    if cpu_steps > 0:      # the checker beside it: CPU oracle on the first cpu_steps steps of the same run
        from oracle.wave_2D_oracle import wave_2D_solve_oracle
        timing = {}
        Ho, _ = wave_2D_solve_oracle(tF * cpu_steps / steps, cpu_steps, mesh, xL, yL, domainShape=shape, K_wave=K, rho=RHO, timing=timing, **kw)
        out['cpu_oracle_steps_per_s'] = cpu_steps / timing['loop_seconds']
        out['H_trace_rel_diff_first_%d_steps' % cpu_steps] = float(np.abs(H[:cpu_steps, 1:] - Ho[:, 1:]).max() / np.abs(Ho[:, 1:]).max())
    return out


def run_paper_bench(args):
    """bench.py --workload paper: one line, same contract; value = all dof-updates of the six cases / their summed device time"""
    res = [run_case(*c, cpu_steps=0 if args.no_cpu_baseline else 200) for c in cases()]
    ms = sum(r['gpu_device_ms'] for r in res)
    upd = sum(r['dofs'] * r['steps'] for r in res)
    line = {'metric': 'dof-updates/sec (fp64, device-timed)', 'value': upd / (ms * 1e-3), 'unit': 'dof-updates/s', 'n_gpus': 1,
            'steps': sum(r['steps'] for r in res), 'warmup': 0, 'ms_per_step': ms / sum(r['steps'] for r in res), 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': 'paper cases (BASELINE configs[0..2]): full-length runs of the six reference configurations through wave_2D_solve, '
                                   'cache-resident meshes (latency-bound; HBM roofline n/a)', 'cases': res},
            'gpu_launches': int(sum(r['launches'] for r in res)),
            'roofline': {'bound': 'hbm', 'achieved': None, 'peak': None, 'unit': 'GB/s', 'frac': None, 'traffic': None,
                         'note': 'cache-resident (SURVEY 8d): steps/s per case is the figure of merit'}}
    if not args.no_cpu_baseline:
        cpu = [r['cpu_oracle_steps_per_s'] * r['dofs'] for r in res]
        line['cpu_baseline'] = {'value': float(np.mean(cpu)), 'unit': 'dof-updates/s', 'cores': 1, 'kind': 'port',
                                'sample': 'first 200 steps of every case with the CPU oracle (scipy SuperLU, factor once), one core; mean over the cases'}
    print(json.dumps(line))


if __name__ == '__main__':
    cpu_steps = int(os.environ.get('CPU_STEPS', '300'))
    todo = cases()
    if os.environ.get('PAPER_ONLY'):        # one case, optionally shortened (ncu captures)
        todo = [todo[int(os.environ['PAPER_ONLY'])]]
    if os.environ.get('PAPER_STEPS'):
        n = int(os.environ['PAPER_STEPS'])
        todo = [c[:5] + (c[5] * n / c[6], n) + c[7:] for c in todo]
    for c in todo:
        print(json.dumps(run_case(*c, cpu_steps=cpu_steps)), flush=True)
