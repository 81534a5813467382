#!/bin/bash
# round 2, GPU call 12: A/B of the clamp-free horizontal path of the cost kernel, of the sweep after the knob clean-up, and of the
# y + C - m add as two multiply-adds (FMA pipe) for 1 / 2 / 4 of the paths
O=gpurun_out; mkdir -p $O
V=build/variants
timeout 900 python tools/ab_variants.py default $V/hfast/libsvgpu.so $V/clean/libsvgpu.so $V/f1/libsvgpu.so $V/f2/libsvgpu.so $V/f4/libsvgpu.so $V/l2a/libsvgpu.so $V/l2b/libsvgpu.so $V/l2c/libsvgpu.so > $O/c12_ab.log 2>&1
cat $O/c12_ab.log
