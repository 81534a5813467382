#!/bin/bash
# Multi-GPU bench lines of the 4K full chain and the 10,000-frame stream (BASELINE configs[3], [4]) on N GPUs of one box.
# usage (through gpurun --gpus N): bash tools/gpu_multi.sh N [configs...]   (default: 4k256_pipeline stream10k)
N="$1"; shift; O=gpurun_out; mkdir -p $O
CFGS="${@:-4k256_pipeline stream10k}"
for cfg in $CFGS; do
  if [ "$N" = 1 ]; then
    timeout 900 python bench.py --gpus 1 --config $cfg --steps 5 --warmup 3 --no-cpu > $O/multi_${cfg}_n$N.json 2> $O/multi_${cfg}_n$N.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --config $cfg --steps 5 --warmup 3 > $O/multi_${cfg}_n$N.json 2> $O/multi_${cfg}_n$N.err
  fi
  python - "$cfg" "$N" "$O/multi_${cfg}_n$N.json" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[3]) if l.startswith('{')][-1])
    print(sys.argv[1], "N =", sys.argv[2], "fps", round(d['details']['fps'], 1), "e2e fps", round(d['e2e']['fps'], 1), "parity", d['details']['parity_checksum'])
except Exception as e:
    print(sys.argv[1], sys.argv[2], "FAILED", e)
PY
  tail -2 $O/multi_${cfg}_n$N.err
done
