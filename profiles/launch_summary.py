#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: mean device time per (kernel, grid).
usage: launch_summary.py launches.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[h]
ki, vi, gi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Grid Size')
d = collections.OrderedDict()
for r in rows[h + 1:]:
    if len(r) > vi:
        try:
            d.setdefault(r[ki][:90] + ' grid=' + r[gi], []).append(float(r[vi].replace(',', '')))
        except ValueError:
            pass
tot = sum(sum(v) for v in d.values())
for k, v in d.items():
    print(f'{sum(v) / len(v) / 1000:8.1f} us x{len(v):3d}  {sum(v) / tot:6.1%}  {k}')
