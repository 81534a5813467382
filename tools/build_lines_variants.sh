#!/bin/bash
# Builds tuning variants of libsvgpu.so in which only sgbm_lines.cu differs: tools/build_lines_variants.sh tag:"-DSV_LINES_PF=8 ..." ...
set -e
cd "$(dirname "$0")/../fhgr-stereo-vision_b200/csrc"
make -s
NV="nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr"
for spec in "$@"; do
  tag="${spec%%:*}"; flags="${spec#*:}"; out=../../build/variants/$tag; mkdir -p $out
  ( $NV $flags -c sgbm_lines.cu -o $out/sgbm_lines.o && nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libsvgpu.so api.o sgbm.o sgbm_cost.o sgbm_sweep.o $out/sgbm_lines.o image_ops.o front_fused.o bm.o rectify.o && rm $out/sgbm_lines.o && echo built $tag ) &
done
wait
