#!/bin/bash
# round 2, GPU call 7: parity suite with the fused front kernel / tiled close / run-based speckle filter; 4K pipeline bench + launch list
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/c7_pytest.log 2>&1; echo "pytest rc=$?" >> $O/c7_pytest.log
tail -15 $O/c7_pytest.log
timeout 900 python bench.py --config 4k256_pipeline --steps 5 --warmup 3 --no-cpu > $O/c7_bench_4k.json 2> $O/c7_bench_4k.err; tail -c 300 $O/c7_bench_4k.json; tail -2 $O/c7_bench_4k.err
python - $O/c7_bench_4k.json <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print("4k pipeline fps", d['config']['fps'], "e2e", d['e2e']['fps'], "ms/step", d['ms_per_step'], d['roofline']['stage_ms_per_step'])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/c7_launches_4k.csv python bench.py --config 4k256_pipeline --steps 1 --warmup 3 --no-cpu > $O/c7_ncu_4k.log 2>&1
python - $O/c7_launches_4k.csv <<'PY'
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]; ki = h.index("Kernel Name"); vi = h.index("Metric Value"); ui = h.index("Metric Unit")
t = collections.Counter(); n = collections.Counter()
for r in rows[1:]:
    v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(r[ui], 1)
    t[r[ki][:50]] += v; n[r[ki][:50]] += 1
tot = sum(t.values())
for k, v in t.most_common(25): print(f"{k:50s} n={n[k]:4d} total {v/1e3:8.2f} ms avg {v/n[k]:9.1f} us {100*v/tot:5.1f}%")
PY
du -sh $O
