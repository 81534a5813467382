#!/bin/bash
# compute-sanitizer memcheck over the small-image parity tests (golden vectors, stages, bottle chain, batched chain): out-of-bounds or
# misaligned accesses of any kernel fail the run.  usage (on the GPU box): bash tools/gpu_sanitize.sh
O=gpurun_out; mkdir -p $O
timeout 1700 compute-sanitizer --tool memcheck --error-exitcode 86 --print-limit 20 \
  python -m pytest tests/test_gpu_parity.py -q -x -k "sgbm_golden_all_modes or stages_golden or bottle_pair_full_chain or pipeline_batch or three_channel or 3way or bm_golden or reconfigure" \
  > $O/sanitize_memcheck.log 2>&1
echo "memcheck rc=$?" >> $O/sanitize_memcheck.log
grep -E "ERROR SUMMARY|Invalid|passed|failed|rc=" $O/sanitize_memcheck.log | tail -12
