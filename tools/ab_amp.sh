#!/bin/bash
# A/B of the three Activation1d kernels inside the full program (fp32 and bf16 paths): tensor-core
# (amp_mma), packed FFMA2 (amp_packed), scalar FFMA.
for prec in fp32 bf16; do
  for cfg in "amp_mma=1" "amp_mma=2 amp_stream=0" "amp_mma=0 amp_stream=0" "amp_mma=0 amp_stream=0 amp_packed=0"; do
    args=""; for kv in $cfg; do args="$args --tune $kv"; done
    echo "== $prec $cfg"; python tools/profile_ops.py --precision $prec $args | tail -17 | grep amp
  done
done
