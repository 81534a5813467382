#!/bin/bash
# The measurement pass of a round on one B200 (through gpurun): parity suite, `ncu --set full` of the hot kernels at 64 pairs
# (-> profiles/r02_ncu_raw.csv, which bench.py reads), the default bench line, its ncu launch list, the other BASELINE
# configurations and the single-call latency table.  Everything lands in gpurun_out/ with the prefix given as $1 (default m).
P="${1:-m}"; O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/${P}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${P}_pytest.log
tail -4 $O/${P}_pytest.log
python tools/ab_variants.py --batch 64 default > $O/${P}_ab.log 2>&1; cat $O/${P}_ab.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_sweep|k_cost_tile" -c 3 -o $O/${P}_hot \
     python tools/ab_variants.py --child --batch 64 --reps 0 --data /tmp/ab_data_1920x1080_128.npz > $O/${P}_ncu_hot.log 2>&1
ncu -i $O/${P}_hot.ncu-rep --page raw --csv > $O/${P}_hot_raw.csv 2>/dev/null
ncu -i $O/${P}_hot.ncu-rep --page source --csv --print-source sass > $O/${P}_hot_src_sass.csv 2>/dev/null; gzip -f $O/${P}_hot_src_sass.csv; rm -f $O/${P}_hot.ncu-rep
cp $O/${P}_hot_raw.csv profiles/r02_ncu_raw.csv
timeout 600 python bench.py --steps 20 --warmup 3 > $O/${P}_bench.json 2> $O/${P}_bench.err; tail -c 400 $O/${P}_bench.json; tail -3 $O/${P}_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/${P}_bench_reference.json 2> $O/${P}_bench_reference.err; tail -c 300 $O/${P}_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/${P}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/${P}_ncu_bench.log 2>&1
for cfg in 720p64 ref_hh4 bm720p 4k256_pipeline stream10k; do
  timeout 900 python bench.py --config $cfg --steps 5 --warmup 3 > $O/${P}_bench_$cfg.json 2> $O/${P}_bench_$cfg.err
  python - "$cfg" "$O/${P}_bench_$cfg.json" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[2]) if l.startswith('{')][-1])
    print(sys.argv[1], "fps", round(d['details']['fps'], 1), "e2e fps", round(d['e2e']['fps'], 1), "cpu", round(d.get('cpu_baseline', {}).get('value', 0)), "value", round(d['value']))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
  tail -2 $O/${P}_bench_$cfg.err
done
timeout 600 python tools/bench_configs.py --latency > $O/${P}_latency.log 2>&1; cat $O/${P}_latency.log
du -sh $O
