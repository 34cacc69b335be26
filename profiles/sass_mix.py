#!/usr/bin/env python
"""Aggregate an `ncu -i X.ncu-rep --page source --csv` dump: executed warp instructions by SASS opcode and
sampled stall reasons.  usage: sass_mix.py dump.csv [n_weight_samples]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if 'Instructions Executed' in r)
hdr = rows[h]
ia, ie, isamp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
stall_cols = [i for i, c in enumerate(hdr) if c.startswith('stall_') and 'Not Issued' not in c]
tot, byop, sampop, stalls = 0, collections.Counter(), collections.Counter(), collections.Counter()
for r in rows[h + 1:]:
    if len(r) <= ie or not r[ie].isdigit():
        continue
    n = int(r[ie]); tot += n
    op = [o for o in r[ia].strip().split() if not o.startswith('@')]
    name = op[0].split('.')[0] if op else '?'
    byop[name] += n; sampop[name] += int(r[isamp] or 0)
    for i in stall_cols:
        stalls[hdr[i]] += int(r[i] or 0)
print('SASS lines', len(rows) - h - 1, 'executed warp instructions', tot)
if len(sys.argv) > 2:
    print('thread instructions per weight-sample', tot * 32 / float(sys.argv[2]))
for k, v in byop.most_common(28):
    print(f'{k:12s} {v:10d} {v / tot:.3f}  samples {sampop[k]}')
print(stalls.most_common(10))
