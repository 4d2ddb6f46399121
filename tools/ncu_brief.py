#!/usr/bin/env python
"""Brief per-kernel digest of an ncu raw-page CSV (ncu -i X.ncu-rep --page raw --csv > raw.csv): duration, instruction count, pipe
utilisation, DRAM bytes and the top stall reasons per issued instruction."""
import csv
import sys

KEYS = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.max',
        'smsp__warps_eligible.avg.per_cycle_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active']


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print('=====', r[idx['Kernel Name']][:60])
        for k in KEYS:
            if k in idx:
                print('  %-72s %s %s' % (k, r[idx[k]], units[idx[k]]))
        vals = []
        for h in hdr:
            if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio'):
                vals.append((float(r[idx[h]]), h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]))
        print('  stalls per issued instruction:', ', '.join('%s %.2f' % (n, v) for v, n in sorted(vals, reverse=True)[:8]))


if __name__ == '__main__':
    main(sys.argv[1])
