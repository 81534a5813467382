#!/bin/bash
# round 2, GPU call 13: parity suite (three-channel shared-ring cost kernel, batch post-filter test), then compute-sanitizer memcheck
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/c13_pytest.log 2>&1; echo "pytest rc=$?" >> $O/c13_pytest.log
tail -12 $O/c13_pytest.log

