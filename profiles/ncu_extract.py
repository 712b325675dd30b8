#!/usr/bin/env python
"""Summarise an `ncu --set full` capture of bench.py (read here, without a GPU: `ncu -i <rep> --page raw --csv`).

    python profiles/ncu_extract.py gpurun_out/r2a_pipe_full.ncu-rep profiles/r2a_ncu_full_pipelined_kernels.txt [--json]

Per launch: duration, DRAM bytes read / written, DRAM throughput (% of pin rate), achieved occupancy, registers, grid, shared-memory
bank conflicts.  For the persistent Chebyshev solves the number of iterations of a launch follows from the bytes written (every
iteration writes the new iterate once: n_rows x 8 B), which turns the launch totals into DRAM bytes per cell and iteration.
--json also writes profiles/ncu_dram_bytes.json, the per-kernel traffic figures bench.py puts into `roofline.traffic`."""
import csv
import io
import json
import os
import subprocess
import sys

N_CELLS, N_V, N_E = 25006592, 12512137, 37518728        # bench mesh (7072 x 1768 x 2)
ALG = {'solve_ell_pipe_k': 40.0, 'solve_mq_pipe_k': 96.0, 'energy_pipe_k': 64.0, 'rhs_edge_pipe_k': 38.0, 'rhs_vertex_pipe_k': 40.0,
       'solve_ell_coop_k': 40.0, 'solve_mq_coop_k': 96.0}
COLS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct']


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {c: hdr.index(c) for c in COLS if c in hdr}
    kname = hdr.index('Kernel Name')

    def val(r, c):
        if c not in ix:
            return float('nan')
        v = float(r[ix[c]].replace(',', ''))
        u = units[ix[c]]
        scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'ms': 1e-3, 'us': 1e-6, 'ns': 1e-9, 's': 1.0, 'usecond': 1e-6, 'msecond': 1e-3, 'nsecond': 1e-9, 'second': 1.0}.get(u)
        return v * scale if scale else v
    lines = ['# {} - per launch (ncu --set full --clock-control none; cold-cache replays, so durations are upper bounds)'.format(os.path.basename(rep)),
             '# kernel | ms | DRAM read GB | DRAM write GB | GB/s moved | DRAM % of pin rate | warps active % | regs | grid x block | smem bank conflicts / wavefronts | iterations | B/cell/iteration moved (algorithmic)']
    agg = {}
    for r in rows[2:]:
        name = r[kname]
        short = name.replace('void ', '').split('<')[0].split('(')[0]
        t, rd, wr = val(r, 'gpu__time_duration.sum'), val(r, 'dram__bytes_read.sum'), val(r, 'dram__bytes_write.sum')
        iters, per = '', ''
        if short.startswith('solve_'):
            nrows = N_V if 'ell' in short else N_E
            n_it = max(1, round(wr / (8.0 * nrows)))
            b = (rd + wr) / n_it / N_CELLS
            iters, per = str(n_it), '{:.1f} ({:.0f})'.format(b, ALG.get(short, 0))
            agg.setdefault(short, []).append(b)
        elif short in ALG:
            b = (rd + wr) / N_CELLS
            per = '{:.1f} ({:.0f})'.format(b, ALG[short])
            agg.setdefault(short, []).append(b)
        lines.append('{} | {:.4f} | {:.3f} | {:.3f} | {:.0f} | {:.1f} | {:.1f} | {:.0f} | {:.0f} x {:.0f} | {:.3g} / {:.3g} | {} | {}'.format(
            short, t * 1e3, rd / 1e9, wr / 1e9, (rd + wr) / t / 1e9, val(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
            val(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'), val(r, 'launch__registers_per_thread'), val(r, 'launch__grid_size'),
            val(r, 'launch__block_size'), val(r, 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'),
            val(r, 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum'), iters, per))
    open(out, 'w').write('\n'.join(lines) + '\n')
    print('\n'.join(lines))
    if '--json' in sys.argv:
        js = {k: {'dram_bytes_per_cell_iteration' if k.startswith('solve_') else 'dram_bytes_per_cell': sum(v) / len(v), 'launches': len(v),
                  'algorithmic': ALG.get(k), 'source': 'dram__bytes_read.sum + dram__bytes_write.sum of the ncu --set full capture {} of this kernel '
                  '(summary: profiles/{}), per cell{}'.format(os.path.basename(rep), os.path.basename(out), ' and Chebyshev iteration, scaled to this launch' if k.startswith('solve_') else '')}
              for k, v in agg.items()}
        json.dump(js, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'ncu_dram_bytes.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
