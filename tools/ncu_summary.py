#!/usr/bin/env python
"""Condenses `ncu -i report.ncu-rep --page raw --csv` into the few counters the
design notes quote (one column per profiled launch).

  ncu -i gpurun_out/x.ncu-rep --page raw --csv > raw.csv
  python tools/ncu_summary.py raw.csv [kernel-name-regex] > profiles/x.txt
"""
import csv
import re
import sys

METRICS = [
    'launch__grid_size', 'launch__block_size', 'gpu__time_duration.sum',
    'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__warps_active.avg.pct_of_peak_sustained_active',
    'launch__registers_per_thread', 'launch__occupancy_limit_registers',
    'launch__occupancy_limit_shared_mem',
    'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'lts__t_sector_hit_rate.pct',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
]


def main():
  rows = list(csv.reader(open(sys.argv[1], newline='')))
  pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
  hdr = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
  names, units = rows[hdr], rows[hdr + 1]
  data = [r for r in rows[hdr + 2:] if len(r) == len(names)]
  kcol = names.index('Kernel Name')
  if pat:
    data = [r for r in data if pat.search(r[kcol])]
  print(f'{"Kernel Name":<92}', [r[kcol][:40] for r in data])
  for m in METRICS:
    if m not in names:
      print('MISSING', m)
      continue
    c = names.index(m)
    print(f'{m:<80} {units[c]:<10}', [r[c] for r in data])


if __name__ == '__main__':
  main()
