#!/usr/bin/env python
"""Source lines ranked by warp-stall samples, from `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`.
usage: stall_lines.py dump.csv [top_n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
cur, out, tot = None, [], 0
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if len(r) > 8 and r[0].isdigit() and r[7].isdigit():
        n, s = int(r[7]), (int(r[6]) if r[6].isdigit() else 0)
        tot += s
        out.append((s, n, cur, int(r[0]), r[1].strip()[:105]))
out.sort(reverse=True)
print('total samples', tot)
for s, n, c, l, src in out[:top]:
    print(f'{s:6d} {s / max(tot, 1):.3f} instr {n:8d}  {c}:{l}  {src}')
