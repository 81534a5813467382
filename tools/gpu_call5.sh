#!/bin/bash
O=gpurun_out; mkdir -p $O
V=build/variants
timeout 900 python tools/ab_variants.py default $V/sr2/libsvgpu.so $V/sr2p/libsvgpu.so $V/sr4/libsvgpu.so $V/sr4p/libsvgpu.so $V/sr2r4/libsvgpu.so > $O/c5_ab.log 2>&1
cat $O/c5_ab.log
