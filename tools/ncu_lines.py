#!/usr/bin/env python
"""Stall samples and executed instructions per CUDA source line from `ncu -i X.ncu-rep --page source --csv --print-source sass,cuda`."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None; hdr = None; acc = collections.Counter(); src = {}; inst = collections.Counter()
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; hdr = None; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and cur and r[0].isdigit():
        try:
            s = int(r[hdr.index("Warp Stall Sampling (All Samples)")] or 0); e = int(r[hdr.index("Instructions Executed")] or 0)
        except Exception:
            continue
        k = (cur, int(r[0])); acc[k] += s; inst[k] += e; src[k] = r[1].strip()[:100]
tot = sum(acc.values()); print('total samples', tot, 'instructions', sum(inst.values()))
for (k, v) in acc.most_common(top):
    print(f"{100*v/tot:5.1f}%  inst {inst[k]/1e6:8.1f}M  {k[0]}:{k[1]:4d}  {src[k]}")
