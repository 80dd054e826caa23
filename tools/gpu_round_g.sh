#!/bin/bash
set -u
OUT=gpurun_out
timeout 300 python tools/conv_precision_diag.py 2>&1 | tail -5
timeout 600 python -m pytest tests -m gpu -q -x > $OUT/r02_pytest_gpu_pair.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/r02_pytest_gpu_pair.log
timeout 300 python tools/precision_diag.py 2>&1 | tail -4
for rep in 1 2; do
  echo "== fp32 pair (default)"; timeout 300 python tools/profile_ops.py --precision fp32 | grep -E "#  conv|total="
  echo "== fp32 umma_pair=0 (stacked single-CTA)"; timeout 300 python tools/profile_ops.py --precision fp32 --tune umma_pair=0 | grep -E "#  conv|total="
done
echo "== bf16 pair"; timeout 300 python tools/profile_ops.py --precision bf16 | grep -E "#  conv|total="
echo "== bf16 umma_pair=0"; timeout 300 python tools/profile_ops.py --precision bf16 --tune umma_pair=0 | grep -E "#  conv|total="
timeout 300 python tools/time_forward.py --parts 2 2>&1 | grep -v Broken | head -6
