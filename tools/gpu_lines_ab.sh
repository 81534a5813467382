#!/bin/bash
# line-kernel variants for single pairs: CFGS / BATCHES select the cases
O=gpurun_out; mkdir -p $O
{
IFS=';'; for cfgs in ${CFGS:-1920x1080 128 1;1280x720 64 0;640x480 224 3}; do
  IFS=' '; set -- $cfgs
  for B in ${BATCHES:-1}; do
    echo "== $1 D=$2 mode=$3 B=$B"
    python tools/ab_variants.py --batch $B --reps 10 --size $1 --disp $2 --mode $3 "${@:4}" $LIBS
  done
done
} 2>&1 | tee $O/lines_ab.txt
