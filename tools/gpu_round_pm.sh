#!/bin/bash
# C = 96 layers (and the time-folded C = 24 ones) on the CTA-pair kernel: parity + per-class times
set -u
OUT=gpurun_out
BVG_TEST_TUNE=umma_pair_min=96 timeout 900 python -m pytest tests/test_gpu_generator.py tests/test_gpu_ops.py -m gpu -q > $OUT/pytest_pm96.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_pm96.log
for P in fp32 bf16; do
  for t in "" "--tune umma_pair_min=96"; do
    echo "== $P $t"
    python tools/profile_ops.py --precision $P $t --out $OUT/ops_pm_tmp.txt > /dev/null 2>&1; grep -E "conv  L(60032|240128|30016|120064)|total=" $OUT/ops_pm_tmp.txt
  done
done
