#!/bin/bash
set -u
for rep in 1 2; do
  echo "== fp32 pair (default)"; timeout 300 python tools/profile_ops.py --precision fp32 | grep -E "#  conv  L(15008|30016|3752)|total="
  echo "== fp32 pair umma_ntile_cap=128"; timeout 300 python tools/profile_ops.py --precision fp32 --tune umma_ntile_cap=128 | grep -E "#  conv  L(15008|30016|3752)|total="
done
echo "== bf16 pair umma_ntile_cap=128"; timeout 300 python tools/profile_ops.py --precision bf16 --tune umma_ntile_cap=128 | grep -E "#  conv  L(15008|30016|3752)|total="
echo "== bf16 pair default"; timeout 300 python tools/profile_ops.py --precision bf16 | grep -E "#  conv  L(15008|30016|3752)|total="
