#!/bin/bash
# round 2, GPU call 1: parity suite, micro-benchmark output, sweep variants, bench, ncu launch list, ncu of the small kernels
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/c1_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x > $O/c1_pytest.log 2>&1; echo "pytest rc=$?" >> $O/c1_pytest.log
tail -3 $O/c1_pytest.log
timeout 120 tools/ubench_int16x2 > $O/r02_ubench_int16x2.txt 2>&1
V=build/variants
timeout 900 python tools/ab_variants.py default $V/i0/libsvgpu.so $V/i1/libsvgpu.so $V/i1em/libsvgpu.so $V/i1mp/libsvgpu.so $V/i1emmp/libsvgpu.so \
   $V/i1emmpf1/libsvgpu.so $V/i1emmpf2/libsvgpu.so $V/i1emmpf4/libsvgpu.so $V/i0emmp/libsvgpu.so > $O/c1_ab.log 2>&1
cat $O/c1_ab.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/c1_bench.json 2> $O/c1_bench.err; tail -c 600 $O/c1_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/c1_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/c1_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_remap|k_gauss|k_sharpen|k_reproject|k_spk|k_morph|k_bm|k_depth|k_min|k_median" -c 40 -o $O/c1_misc python tools/run_misc_kernels.py > $O/c1_ncu_misc.log 2>&1
tail -2 $O/c1_ncu_misc.log
