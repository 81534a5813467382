#!/bin/bash
# round 2, GPU call 3: source-level profiles (exported to CSV on the box) of the cost kernel and of the first sweep with / without
# the interior copy; A/B of the interior knobs
O=gpurun_out; mkdir -p $O
V=build/variants
python tools/ab_variants.py --batch 64 default > $O/c3_warm.log 2>&1   # creates the cached inputs
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_cost_tile" -c 1 -o $O/c3_cost \
     python tools/ab_variants.py --child --batch 16 --reps 0 --data /tmp/ab_data_1920x1080_128.npz > $O/c3_ncu_cost.log 2>&1
tools/ncu_export.sh $O/c3_cost.ncu-rep
for v in k0 k1; do
  SVGPU_LIB=$PWD/$V/$v/libsvgpu.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_sweep" -c 1 -o $O/c3_sweep_$v \
     python tools/ab_variants.py --child --batch 16 --reps 0 --data /tmp/ab_data_1920x1080_128.npz > $O/c3_ncu_$v.log 2>&1
  tools/ncu_export.sh $O/c3_sweep_$v.ncu-rep
done
timeout 900 python tools/ab_variants.py default $V/k0/libsvgpu.so $V/k1/libsvgpu.so $V/k3/libsvgpu.so $V/k5/libsvgpu.so $V/k9/libsvgpu.so $V/k17/libsvgpu.so \
   $V/k27/libsvgpu.so $V/k29/libsvgpu.so $V/k23/libsvgpu.so $V/k1a/libsvgpu.so $V/k0a/libsvgpu.so > $O/c3_ab.log 2>&1
cat $O/c3_ab.log
du -sh $O
