#!/bin/bash
set -u
for c in 0 7 6 5 4; do
  echo "== fp32 amp_chunk=$c"; timeout 300 python tools/profile_ops.py --precision fp32 --tune amp_chunk=$c | grep -E "#  amp"
done
echo "== fp32 amp_chunk=0 again"; timeout 300 python tools/profile_ops.py --precision fp32 | grep -E "#  amp"
