#!/bin/bash
# Builds tuning variants of libsvgpu.so (only sgbm_sweep.cu differs) into build/variants/<tag>/libsvgpu.so.
# usage: tools/build_variants.sh tag1:"-DSV_SWEEP_K=8 ..." tag2:"..."   then run with SVGPU_LIB=build/variants/<tag>/libsvgpu.so
set -e
cd "$(dirname "$0")/../fhgr-stereo-vision_b200/csrc"
make -s
for spec in "$@"; do
  tag="${spec%%:*}"; flags="${spec#*:}"
  out="../../build/variants/$tag"; mkdir -p "$out"
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr $flags -c sgbm_sweep.cu -o "$out/sgbm_sweep.o" &
done
wait
for spec in "$@"; do
  tag="${spec%%:*}"; out="../../build/variants/$tag"
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$out/libsvgpu.so" api.o sgbm.o sgbm_cost.o "$out/sgbm_sweep.o" sgbm_lines.o image_ops.o front_fused.o bm.o rectify.o
  echo "built $out/libsvgpu.so"
done
