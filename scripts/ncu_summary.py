#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): headline metrics + stall reasons per kernel."""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed.sum', 'smsp__inst_executed.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_blocks',
        'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
        'launch__grid_size', 'launch__block_size', 'smsp__cycles_active.avg',
        'sm__cycles_elapsed.avg.per_second']


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('###', r[hdr.index('Kernel Name')][:100])
        for w in WANT:
            if w in hdr:
                print('  %-62s %16s %s' % (w, r[hdr.index(w)], units[hdr.index(w)]))
        stalls = [(float(r[i]), h) for i, h in enumerate(hdr)
                  if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')
                  and r[i] not in ('', 'n/a')]
        for v, h in sorted(stalls, reverse=True)[:8]:
            print('  stall %-56s %8.3f' % (h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v))


if __name__ == '__main__':
    main(sys.argv[1])
