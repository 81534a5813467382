#!/bin/bash
# A/B run of the tuning variants built by tools/build_variants.sh: prints fps and stage times per variant and batch.
# usage (on the GPU box): tools/ab_bench.sh "8 16" tag1 tag2 ...   ("default" = the in-tree build)
batches="$1"; shift
for tag in "$@"; do
  for b in $batches; do
    lib=""; [ "$tag" != default ] && lib="$PWD/build/variants/$tag/libsvgpu.so"
    SVGPU_LIB="$lib" timeout 200 python bench.py --steps 5 --warmup 3 --batch $b --no-cpu > /tmp/ab.json 2> /tmp/ab.err
    python - "$tag" "$b" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open('/tmp/ab.json') if l.startswith('{')][-1])
    s = d['roofline']['stage_ms_per_step']
    print(f"{sys.argv[1]:12s} B={sys.argv[2]:3s} fps {d['config']['fps']:7.1f} e2e {d['e2e']['fps']:7.1f} cost {s['cost']:6.2f} agg {s['aggregate']:6.2f} sel {s['select']:5.2f}", flush=True)
except Exception as e:
    print(sys.argv[1], sys.argv[2], 'FAILED', e, open('/tmp/ab.err').read()[-300:])
PY
  done
done
