#!/bin/bash
set -u
for rep in 1 2; do
  echo "== fp32 default (stacked 128 for C>=384)"; python tools/profile_ops.py --precision fp32 | grep "#  conv"
  echo "== fp32 umma_stack=0 (256/192-column tiles, 3 separate passes)"; python tools/profile_ops.py --precision fp32 --tune umma_stack=0 | grep "#  conv"
done
echo "== bf16 default"; python tools/profile_ops.py --precision bf16 | grep "#  conv"
echo "== bf16 umma_ntile_cap=128"; python tools/profile_ops.py --precision bf16 --tune umma_ntile_cap=128 | grep "#  conv"
echo "== bf16 default"; python tools/profile_ops.py --precision bf16 | grep "#  conv"
