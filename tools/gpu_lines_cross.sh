#!/bin/bash
# crossover between the line form and the strip pipeline (forced either way) by pairs per call
O=gpurun_out; mkdir -p $O
run() { # size disp mode batches...
  local size=$1 disp=$2 mode=$3; shift 3
  for B in "$@"; do
    for LN in 0 1; do
      echo -n "lines=$LN $size D=$disp mode=$mode "
      SVGPU_SWEEP_LINES=$LN python tools/ab_variants.py --batch $B --reps 6 --size $size --disp $disp --mode $mode default
    done
  done
}
{
run 1920x1080 128 1 1 2 3
run 1280x720 64 0 2 4 8 12
run 640x480 224 3 2 4 8 12
} 2>&1 | tee $O/lines_cross.txt
