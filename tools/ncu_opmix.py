#!/usr/bin/env python
"""Dynamic SASS opcode mix of one kernel from `ncu -i X.ncu-rep --page source --csv` (developer tool).
usage: ncu_opmix.py source.csv [divisor [kernel-index]]   -- counts are divided by `divisor` (e.g. warp-steps of the launch)"""
import csv, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = rows[1]
iS, iE, iSm = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
ops, samp, tot, totS = Counter(), Counter(), 0, 0
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
k = -1
for r in rows:
    if r and r[0] == 'Kernel Name':
        k += 1
        continue
    if k != which or len(r) <= max(iS, iE, iSm) or not r[iE].isdigit():
        continue
    s = r[iS].strip().split()
    if not s:
        continue
    op = s[1] if s[0].startswith('@') and len(s) > 1 else s[0]
    op = op.rstrip(';')
    key = '.'.join(op.split('.')[:2]) if op.split('.')[0] in ('IMAD', 'LDS', 'STS', 'LDG', 'STG', 'SHFL') else op.split('.')[0]
    e = int(r[iE] or 0); ops[key] += e; tot += e
    sm = int(r[iSm] or 0); samp[key] += sm; totS += sm
for k, v in ops.most_common(40):
    print(f"{k:14s} {v / div:10.1f}  {100 * v / tot:5.1f}%  stall-samples {100 * samp[k] / max(totS, 1):5.1f}%")
print('total', tot / div)
