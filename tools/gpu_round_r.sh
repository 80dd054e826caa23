#!/bin/bash
set -u
OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/r02_pytest_gpu_epi.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/r02_pytest_gpu_epi.log
timeout 300 python tools/profile_ops.py --precision bf16 | grep -E "#  conv|total="
timeout 300 python tools/profile_ops.py --precision fp32 | grep -E "#  conv  L(60032|120064|240128)|total="
timeout 300 python tools/time_forward.py --parts 0,2 2>&1 | grep -v Broken | head -10
