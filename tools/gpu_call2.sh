#!/bin/bash
# round 2, GPU call 2: parity suite (reproject fix, batch pipeline), cost-kernel variants, sweep interior experiments + source-level ncu
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/c2_pytest.log 2>&1; echo "pytest rc=$?" >> $O/c2_pytest.log
tail -5 $O/c2_pytest.log
V=build/variants
timeout 900 python tools/ab_variants.py default $V/c2m1/libsvgpu.so $V/c2m2/libsvgpu.so $V/c2m1r8/libsvgpu.so $V/c2n128/libsvgpu.so \
   $V/j0/libsvgpu.so $V/j1/libsvgpu.so $V/j1f/libsvgpu.so $V/j0f/libsvgpu.so > $O/c2_ab.log 2>&1
cat $O/c2_ab.log
# source-level profiles of the first sweep, 16 pairs per launch: interior copy off / on
for v in j0 j1; do
  SVGPU_LIB=$PWD/$V/$v/libsvgpu.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_sweep" -c 2 -o $O/c2_sweep_$v \
     python tools/ab_variants.py --child --batch 16 --reps 0 --data /tmp/ab_data_1920x1080_128.npz > $O/c2_ncu_$v.log 2>&1
  tail -1 $O/c2_ncu_$v.log
done
SVGPU_LIB=$PWD/$V/c2m1/libsvgpu.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_cost_tile" -c 1 -o $O/c2_cost_c2m1 \
     python tools/ab_variants.py --child --batch 16 --reps 0 --data /tmp/ab_data_1920x1080_128.npz > $O/c2_ncu_cost.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_cost_tile" -c 1 -o $O/c2_cost_c1 \
     python tools/ab_variants.py --child --batch 16 --reps 0 --data /tmp/ab_data_1920x1080_128.npz > $O/c2_ncu_cost1.log 2>&1
ls -la $O/*.ncu-rep
