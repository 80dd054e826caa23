#!/bin/bash
set -u
OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $OUT/r02_pytest_gpu_pair3.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/r02_pytest_gpu_pair3.log
grep -E "diffsvc step|recipe " $OUT/r02_pytest_gpu_pair3.log | head
timeout 300 python tools/precision_diag.py 2>&1 | tail -3 | cut -c1-170
timeout 300 python tools/time_forward.py --parts 2 2>&1 | grep -v Broken | head -6
