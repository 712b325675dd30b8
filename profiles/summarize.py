"""Summaries of the ncu captures kept under profiles/ (run here, no GPU needed):
   python profiles/summarize.py launches <csv>     -> per-kernel totals / shares of a --metrics gpu__time_duration.sum run
   python profiles/summarize.py full <ncu-rep>     -> key metrics of a --set full capture
"""
import collections
import csv
import re
import subprocess
import sys


def launches(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = re.sub(r'\(.*', '', row['Kernel Name']).strip()
        v = float(row['Metric Value'].replace(',', ''))
        unit = row['Metric Unit']
        v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'msecond': 1.0, 'usecond': 1e-3, 'nsecond': 1e-6, 's': 1e3}.get(unit, 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print('{:>10} {:>8} {:>12} {:>7}  kernel'.format('total ms', 'launches', 'ms/launch', 'share'))
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('{:10.3f} {:8d} {:12.4f} {:6.1f}%  {}'.format(t, n, t / n, 100 * t / tot, k[:100]))


WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum.per_second', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fp64.sum']


def full(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print('---- ' + r[idx['Kernel Name']])
        for w in WANT:
            if w in idx:
                print('  {:72s} {:>18s} {}'.format(w, r[idx[w]], units[idx[w]]))
        stalls = [(float(r[i].replace(',', '')), h) for h, i in idx.items() if h.startswith('smsp__average_warp') and 'issue_stalled' in h and h.endswith('_per_issue_active.ratio') and r[i]]
        for v, h in sorted(stalls, reverse=True)[:6]:
            print('  stall {:66s} {:18.3f}'.format(h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v))


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2])
