#!/bin/bash
# Quick experiment builds: only the D = 128 MODE_HH throughput kernels of the sweep (-DSV_SWEEP_ONLY8) and/or the cost kernel
# are recompiled with the given flags; everything else is the in-tree build.
# usage: tools/build_fast_variants.sh tag1:"-DSV_SWEEP_EDGEMAD=1" tag2:"cost:-DSV_COST_X=1" ...   (prefix "cost:" = flags for sgbm_cost.cu)
set -e
cd "$(dirname "$0")/../fhgr-stereo-vision_b200/csrc"
NV="nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr"
for spec in "$@"; do
  tag="${spec%%:*}"; flags="${spec#*:}"
  out="../../build/variants/$tag"; mkdir -p "$out"
  if [[ "$flags" == cost:* ]]; then
    ( $NV ${flags#cost:} -c sgbm_cost.cu -o "$out/sgbm_cost.o" 2> "$out/build.log" && cp sgbm_sweep.o "$out/sgbm_sweep.o" ) &
  else
    ( $NV -DSV_SWEEP_ONLY8 $flags -c sgbm_sweep.cu -o "$out/sgbm_sweep.o" 2> "$out/build.log" && cp sgbm_cost.o "$out/sgbm_cost.o" ) &
  fi
  while [ "$(jobs -r | wc -l)" -ge 8 ]; do sleep 0.5; done
done
wait
for spec in "$@"; do
  tag="${spec%%:*}"; out="../../build/variants/$tag"
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$out/libsvgpu.so" api.o sgbm.o "$out/sgbm_cost.o" "$out/sgbm_sweep.o" sgbm_lines.o image_ops.o front_fused.o bm.o rectify.o
  if [ -z "$KEEP_OBJ" ]; then rm -f "$out/sgbm_cost.o" "$out/sgbm_sweep.o"; fi   # the snapshot sent to the GPU box is limited to 512 MiB
  echo "built $out/libsvgpu.so"
done
