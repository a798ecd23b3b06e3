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
    lines = [l for l in open(path) if not l.startswith('==')]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    n = 0
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(row['Metric Value'].replace(',', ''))
        v = {'ns': v / 1e3, 'us': v, 'ms': v * 1e3, 's': v * 1e6}[row['Metric Unit']]
        k = re.sub(r'\(anonymous namespace\)::|<unnamed>::', '', row['Kernel Name'])
        k = re.sub(r'^void ', '', k)
        k = re.sub(r'\(.*', '', k)[:100]
        tot[k] += v
        cnt[k] += 1
        n += 1
    T = sum(tot.values())
    print(f'{n} launches, {T / 1e3:.2f} ms total device time (ncu: serialised, cold cache -- compare SHARES)')
    for k, v in sorted(tot.items(), key=lambda x: -x[1])[:40]:
        print(f'{v / 1e3:10.3f} ms {100 * v / T:6.2f}%  x{cnt[k]:5d}  avg {v / cnt[k]:10.1f} us  {k}')


if __name__ == '__main__':
    {'kernel': kernel, 'launches': launches}[sys.argv[1]](sys.argv[2])
