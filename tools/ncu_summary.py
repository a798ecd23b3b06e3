#!/usr/bin/env python
"""Turn ncu outputs under gpurun_out/ into the small text summaries committed under profiles/.
  python tools/ncu_summary.py launches gpurun_out/launches_r01.csv > profiles/r01_launches.txt
  python tools/ncu_summary.py kernel gpurun_out/sca_fwd_r01.ncu-rep > profiles/r01_sca_fwd.txt"""
import collections
import csv
import re
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'lts__t_bytes.sum', 'lts__t_sectors_op_atom.sum', 'lts__t_sectors_op_red.sum']
STALLS = 'smsp__average_warps_issue_stalled_'


def kernel(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')]
        print(f'kernel: {name[:150]}')
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f'  {w:78s} {r[i]:>18s} {units[i]}')
        stalls = [(float(r[i] or 0), h[len(STALLS):-len('_per_issue_active.ratio')]) for i, h in enumerate(hdr)
                  if h.startswith(STALLS) and h.endswith('_per_issue_active.ratio')]
        print('  top stall reasons (warps per issue-active): ' +
              ', '.join(f'{n}={v:.2f}' for v, n in sorted(stalls, reverse=True)[:6]))
        print()


def launches(path):
    """per kernel name: device time (share), and -- when the capture has them -- the time-weighted tensor-pipe
    activity and the DRAM bytes per launch"""
    lines = [l for l in open(path) if not l.startswith('==')]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    tens, dram = collections.defaultdict(float), collections.defaultdict(float)
    per_id = collections.defaultdict(dict)
    for row in csv.DictReader(lines):
        k = re.sub(r'\(anonymous namespace\)::|<unnamed>::', '', row['Kernel Name'])
        k = re.sub(r'^void ', '', k)
        k = re.sub(r'\(.*', '', k)[:100]
        v = float(row['Metric Value'].replace(',', ''))
        per_id[(row['ID'], k)][row['Metric Name']] = (v, row['Metric Unit'])
    n = 0
    for (_, k), m in per_id.items():
        if 'gpu__time_duration.sum' not in m:
            continue
        v, unit = m['gpu__time_duration.sum']
        us = {'ns': v / 1e3, 'us': v, 'ms': v * 1e3, 's': v * 1e6}[unit]
        tot[k] += us
        cnt[k] += 1
        n += 1
        t = m.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')
        if t:
            tens[k] += t[0] * us
        for name in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
            if name in m:
                b, bu = m[name]
                dram[k] += b * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[bu]
    T = sum(tot.values())
    print(f'{n} launches, {T / 1e3:.2f} ms total device time (ncu: serialised, cold cache -- compare SHARES)')
    print('   time      share   launches   avg us   tensor-pipe %   DRAM MB/launch   kernel')
    for k, v in sorted(tot.items(), key=lambda x: -x[1])[:45]:
        print(f'{v / 1e3:8.3f} ms {100 * v / T:6.2f}%  x{cnt[k]:4d} {v / cnt[k]:9.1f} {tens[k] / v if v else 0:10.1f} '
              f'{dram[k] / cnt[k] / 1e6:14.1f}     {k}')


if __name__ == '__main__':
    {'kernel': kernel, 'launches': launches}[sys.argv[1]](sys.argv[2])
