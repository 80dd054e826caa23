#!/bin/bash
set -u
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k conv 2>&1 | tail -2
for rep in 1 2; do
  echo "== fp32 pair default"; timeout 300 python tools/profile_ops.py --precision fp32 | grep -E "#  conv  L(15008|30016|3752|938)|total="
  echo "== fp32 umma_pair=3 (128-column split pair tiles, two stages)"; timeout 300 python tools/profile_ops.py --precision fp32 --tune umma_pair=3 | grep -E "#  conv  L(15008|30016|3752|938)|total="
done
timeout 300 python tools/time_forward.py --precisions fp32 --parts 2 2>&1 | grep -v Broken | head -3
timeout 300 python tools/time_forward.py --precisions fp32 --parts 2 --tune umma_pair=3 2>&1 | grep -v Broken | head -3
timeout 300 python tools/time_forward.py --precisions fp32 --parts 2 2>&1 | grep -v Broken | head -3
timeout 300 python tools/time_forward.py --precisions fp32 --parts 2 --tune umma_pair=3 2>&1 | grep -v Broken | head -3
