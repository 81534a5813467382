#!/bin/bash
O=gpurun_out; mkdir -p $O
V=build/variants
timeout 900 python tools/ab_variants.py default $V/c2n128/libsvgpu.so $V/sr1/libsvgpu.so $V/sr2/libsvgpu.so $V/sr2b/libsvgpu.so > $O/c4_ab.log 2>&1
cat $O/c4_ab.log
