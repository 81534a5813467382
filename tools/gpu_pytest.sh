#!/bin/bash
# The GPU parity suite on the box: bash tools/gpu_pytest.sh   (through gpurun)
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -12 $O/pytest_gpu.log
