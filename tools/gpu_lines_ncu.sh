#!/bin/bash
# ncu --set full of the line kernel of one single-pair call: tools/gpu_lines_ncu.sh "1280x720 64 0" [lib]
O=gpurun_out; mkdir -p $O
set -- $1 $2
python tools/ab_variants.py --batch 1 --reps 1 --size $1 --disp $2 --mode $3 default > /dev/null 2>&1   # creates the cached inputs
[ -n "$4" ] && export SVGPU_LIB=$PWD/$4
ncu --set full --import-source on --clock-control none -k regex:"k_lines" --launch-skip 2 -c 1 -o $O/lines \
  python tools/ab_variants.py --child --batch 1 --reps 1 --size $1 --disp $2 --mode $3 --data /tmp/ab_data_$1_$2.npz > $O/lines_ncu.log 2>&1
tail -2 $O/lines_ncu.log
bash tools/ncu_export.sh $O/lines.ncu-rep
