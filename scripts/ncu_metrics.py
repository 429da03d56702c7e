#!/usr/bin/env python
"""Print selected raw metrics of an .ncu-rep: ncu_metrics.py file.ncu-rep substr [substr...]"""
import csv, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print('###', r[hdr.index('Kernel Name')][:90])
    for h, u, v in zip(hdr, units, r):
        if any(k in h for k in sys.argv[2:]):
            print('  %-90s %14s %s' % (h, v, u))
