#!/bin/bash
# round 2, GPU call 10: cost kernel with TMA row staging vs without, then the parity suite
O=gpurun_out; mkdir -p $O
V=build/variants
timeout 900 python tools/ab_variants.py default $V/notma/libsvgpu.so > $O/c10_ab.log 2>&1
cat $O/c10_ab.log
timeout 1500 python -m pytest tests -m gpu -q -x > $O/c10_pytest.log 2>&1; echo "pytest rc=$?" >> $O/c10_pytest.log
tail -8 $O/c10_pytest.log
