#!/usr/bin/env python
"""Per-source-line executed warp instructions from `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`.
usage: line_mix.py dump.csv [top_n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, out, tot = None, [], 0
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if len(r) > 8 and r[0].isdigit() and r[7].isdigit():      # a source line row: Line No, Source, ..., # Samples, Instr Exec
        n = int(r[7]); tot += n
        out.append((n, int(r[6]) if r[6].isdigit() else 0, cur, int(r[0]), r[1].strip()[:110]))
out.sort(reverse=True)
print('total executed warp instructions', tot)
for n, samp, f, ln, src in out[:top]:
    print(f'{n:9d} {n / tot:6.3f} samp {samp:5d}  {f}:{ln}  {src}')
