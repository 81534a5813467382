#!/usr/bin/env python
"""Top SASS instructions by warp-stall samples from `ncu -i X.ncu-rep --page source --csv` (developer tool).
usage: ncu_topstalls.py source.csv [N [kernel-index]]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hdr = rows[1]
iS, iE, iSm = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
body, k = [], -1
for r in rows:
    if r and r[0] == 'Kernel Name':
        k += 1
        continue
    if k != which or len(r) <= iSm or not r[iE].isdigit():
        continue
    body.append((int(r[iSm]), int(r[iE]), r[iS].strip(), len(body)))
tot = sum(b[0] for b in body)
print('total samples', tot, 'warp instructions', sum(b[1] for b in body))
for s, e, src, i in sorted(body, key=lambda t: -t[0])[:N]:
    print(f"{i:5d} {100 * s / tot:5.1f}% exec {e:10d}  {src[:100]}")
