#!/bin/bash
O=gpurun_out; mkdir -p $O
V=build/variants
timeout 900 python tools/ab_variants.py default $V/p0/libsvgpu.so $V/p1/libsvgpu.so $V/p2/libsvgpu.so > $O/c8_ab.log 2>&1
cat $O/c8_ab.log
