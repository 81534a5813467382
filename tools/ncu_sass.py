#!/usr/bin/env python
"""Stall samples per SASS instruction from `ncu -i X --page source --csv --print-source sass` (tools/ncu_export.sh):
totals per stall reason, per opcode, and the hottest instructions with their neighbourhood.
usage: tools/ncu_sass.py file.csv [top] [context]"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
I = {n: i for i, n in enumerate(hdr)}
reasons = [n for n in hdr if n.startswith("stall_") and "(Not Issued)" not in n]
data = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    data.append(r)
tot = sum(int(r[I["# Samples"]] or 0) for r in data)
inst = sum(int(r[I["Instructions Executed"]] or 0) for r in data)
print(f"kernel: {rows[0][1][:100]}\ninstructions (static) {len(data)}, executed {inst/1e6:.1f} M warp-instr, samples {tot}")
print("by reason:", ", ".join(f"{n[6:]} {100*sum(int(r[I[n]] or 0) for r in data)/tot:.1f}%" for n in sorted(reasons, key=lambda n: -sum(int(r[I[n]] or 0) for r in data))[:10]))
ops = collections.Counter(); opi = collections.Counter()
for r in data:
    src = re.sub(r"^@!?U?P\w+\s+", "", r[I["Source"]].strip())
    op = src.split()[0].split(".")[0] if src else "?"
    ops[op] += int(r[I["# Samples"]] or 0); opi[op] += int(r[I["Instructions Executed"]] or 0)
print("by opcode (samples %, executed %):", ", ".join(f"{o} {100*v/tot:.1f}/{100*opi[o]/inst:.1f}" for o, v in ops.most_common(16)))
order = sorted(range(len(data)), key=lambda i: -int(data[i][I["# Samples"]] or 0))[:top]
for i in order:
    r = data[i]
    rs = sorted(((int(r[I[n]] or 0), n[6:]) for n in reasons), reverse=True)[:3]
    print(f"{100*int(r[I['# Samples']])/tot:5.2f}%  #{i:5d} exec {int(r[I['Instructions Executed']])/1e6:7.2f}M  {r[I['Source']].strip()[:70]:70s} " + " ".join(f"{n}:{v}" for v, n in rs if v))
