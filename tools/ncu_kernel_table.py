#!/usr/bin/env python
"""Per-kernel table from `ncu -i X.ncu-rep --page raw --csv`: time, DRAM bytes and fraction of the measured HBM peak, pipes, stalls.
usage: ncu -i rep --page raw --csv | python tools/ncu_kernel_table.py [hbm_GBs]"""
import csv, sys, collections
rows = list(csv.reader(sys.stdin))
h = rows[0]
units = rows[1]
SCALE = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
peak = float(sys.argv[1]) if len(sys.argv) > 1 else 6470.5
def g(r, name, d=0.0):
    try:
        i = h.index(name)
        return float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0)
    except Exception:
        return d
agg = collections.OrderedDict()
for r in rows[2:]:
    k = r[h.index("Kernel Name")]
    agg.setdefault(k, []).append(r)
print("| kernel | launches | us | DRAM rd MB | DRAM wr MB | GB/s | frac of HBM | ALU % | FMA % | LSU % | issue % | regs | warps/sched |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for k, rs in agg.items():
    r = rs[-1]
    t = g(r, "gpu__time_duration.sum") * 1e6
    rd, wr = g(r, "dram__bytes_read.sum"), g(r, "dram__bytes_write.sum")
    bw = (rd + wr) / (t * 1e-6) / 1e9 if t else 0
    print(f"| {k[:60]} | {len(rs)} | {t:.1f} | {rd/1e6:.1f} | {wr/1e6:.1f} | {bw:.0f} | {bw/peak:.2f} | "
          f"{g(r,'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'):.0f} | {g(r,'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active'):.0f} | "
          f"{g(r,'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active'):.0f} | {g(r,'sm__inst_issued.avg.pct_of_peak_sustained_active'):.0f} | "
          f"{g(r,'launch__registers_per_thread'):.0f} | {g(r,'sm__warps_active.avg.per_cycle_active')/4:.1f} |")
