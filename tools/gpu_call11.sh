#!/bin/bash
# round 2, GPU call 11: SASS-level profile of the cost kernel with TMA staging (64 pairs), exported to CSV on the box
O=gpurun_out; mkdir -p $O
python tools/ab_variants.py --batch 64 default > $O/c11_ab.log 2>&1; cat $O/c11_ab.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_cost_tile" -c 1 -o $O/c11_cost \
     python tools/ab_variants.py --child --batch 64 --reps 0 --data /tmp/ab_data_1920x1080_128.npz > $O/c11_ncu.log 2>&1
tools/ncu_export.sh $O/c11_cost.ncu-rep
