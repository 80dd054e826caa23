#!/bin/bash
set -u
OUT=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > $OUT/r02_pytest_gpu_pair2.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/r02_pytest_gpu_pair2.log
for rep in 1 2; do
  echo "== fp32 default (C=192 on the pair kernel, one accumulator, two stages)"; timeout 300 python tools/profile_ops.py --precision fp32 | grep -E "#  conv  L(15008|30016|3752)|total="
  echo "== fp32 umma_pair=2 (C=192 on the single-CTA kernel)"; timeout 300 python tools/profile_ops.py --precision fp32 --tune umma_pair=2 | grep -E "#  conv  L(15008|30016|3752)|total="
done
timeout 300 python tools/precision_diag.py 2>&1 | tail -3 | cut -c1-150
