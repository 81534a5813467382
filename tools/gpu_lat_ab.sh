#!/bin/bash
# A/B of sweep variants at several pairs per call (1080p/128 MODE_HH): BATCHES="1 2 4 8" tools/gpu_lat_ab.sh lib...
O=gpurun_out; mkdir -p $O
for B in ${BATCHES:-1 2 4 8}; do
  python tools/ab_variants.py --batch $B --reps ${REPS:-10} ${ABARGS:-} "$@"
done 2>&1 | tee $O/lat_ab.txt
