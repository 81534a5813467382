#!/bin/bash
# parity suite + the configurations with more than 128 disparities after the ring-depth change of the 16-lanes-per-pixel sweeps
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/l16_pytest.log 2>&1; echo "pytest rc=$?" >> $O/l16_pytest.log
tail -6 $O/l16_pytest.log
python tools/ab_variants.py --batch 64 default 2>&1 | tail -2
python tools/ab_variants.py --batch 8 --size 3840x2160 --disp 256 default 2>&1 | tail -2
python tools/ab_variants.py --batch 64 --size 640x480 --disp 224 --mode 3 default 2>&1 | tail -2
for cfg in 4k256_pipeline ref_hh4; do
  timeout 900 python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu > $O/l16_bench_$cfg.json 2> $O/l16_bench_$cfg.err
  python - "$cfg" "$O/l16_bench_$cfg.json" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[2]) if l.startswith('{')][-1])
    print(sys.argv[1], "fps", round(d['config']['fps'], 1), "e2e fps", round(d['e2e']['fps'], 1), d['roofline']['stage_ms_per_step'])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
