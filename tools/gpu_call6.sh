#!/bin/bash
# round 2, GPU call 6: parity suite with the two-column cost kernel, ncu of the hot kernels at 64 pairs (-> profiles/r02_ncu_raw.csv),
# bench default + launch list, the other bench configs, single-call latency table
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/c6_pytest.log 2>&1; echo "pytest rc=$?" >> $O/c6_pytest.log
tail -4 $O/c6_pytest.log
python tools/ab_variants.py --batch 64 default > $O/c6_ab.log 2>&1; cat $O/c6_ab.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_sweep|k_cost_tile" -c 3 -o $O/c6_hot \
     python tools/ab_variants.py --child --batch 64 --reps 0 --data /tmp/ab_data_1920x1080_128.npz > $O/c6_ncu_hot.log 2>&1
ncu -i $O/c6_hot.ncu-rep --page raw --csv > $O/c6_hot_raw.csv 2>/dev/null
ncu -i $O/c6_hot.ncu-rep --page source --csv --print-source sass > $O/c6_hot_src_sass.csv 2>/dev/null; gzip -f $O/c6_hot_src_sass.csv; rm -f $O/c6_hot.ncu-rep
cp $O/c6_hot_raw.csv profiles/r02_ncu_raw.csv
timeout 600 python bench.py --steps 10 --warmup 3 > $O/c6_bench.json 2> $O/c6_bench.err; tail -c 400 $O/c6_bench.json; tail -3 $O/c6_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/c6_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/c6_ncu_bench.log 2>&1
for cfg in 720p64 ref_hh4 bm720p 4k256_pipeline stream10k; do
  timeout 900 python bench.py --config $cfg --steps 5 --warmup 3 > $O/c6_bench_$cfg.json 2> $O/c6_bench_$cfg.err
  python - "$cfg" "$O/c6_bench_$cfg.json" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[2]) if l.startswith('{')][-1])
    print(sys.argv[1], "fps", round(d['config']['fps'], 1), "e2e fps", round(d['e2e']['fps'], 1), "cpu", round(d.get('cpu_baseline', {}).get('value', 0)), "value", round(d['value']))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
  tail -2 $O/c6_bench_$cfg.err
done
timeout 600 python tools/bench_configs.py --latency > $O/c6_latency.log 2>&1; cat $O/c6_latency.log
du -sh $O
