#!/bin/bash
set -u
for rep in 1 2; do
  echo "== bf16 default smem"; timeout 300 python tools/time_forward.py --precisions bf16 --parts 0,2 2>&1 | grep -v Broken | head -4
  echo "== bf16 umma_pair_smem_kb=150"; timeout 300 python tools/time_forward.py --precisions bf16 --parts 0,2 --tune umma_pair_smem_kb=150 2>&1 | grep -v Broken | head -4
  echo "== bf16 umma_pair_smem_kb=120"; timeout 300 python tools/time_forward.py --precisions bf16 --parts 0,2 --tune umma_pair_smem_kb=120 2>&1 | grep -v Broken | head -4
done
