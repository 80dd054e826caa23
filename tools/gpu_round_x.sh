#!/bin/bash
# DiffSVC step: tile-size knobs at larger row counts
for shape in "4 379" "16 938"; do
  set -- $shape
  for prec in fp32 bf16; do
    for t in "--tune umma_ntile_cap=32" "--tune umma_ntile_cap=64" "--tune umma_ntile_cap=128" "--tune umma_ntile_cap=256"; do
      python tools/profile_diffsvc.py --precision $prec --eager 0 --brief 1 --batch $1 --frames $2 $t 2>&1 | grep -E "forward"
    done
  done
done
python tools/profile_diffsvc.py --precision fp32 --eager 0 --brief 1 --tune umma_ntile_cap=16 2>&1 | grep -E "forward"
python tools/profile_diffsvc.py --precision bf16 --eager 0 --brief 1 --tune umma_ntile_cap=16 2>&1 | grep -E "forward"
