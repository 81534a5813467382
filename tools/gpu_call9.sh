#!/bin/bash
O=gpurun_out; mkdir -p $O
V=build/variants
timeout 900 python tools/ab_variants.py default $V/p3/libsvgpu.so $V/p4/libsvgpu.so > $O/c9_ab.log 2>&1
cat $O/c9_ab.log
