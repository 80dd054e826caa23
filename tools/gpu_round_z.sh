#!/bin/bash
# vocoder single-utterance latency vs tile width
for t in "" "--tune umma_ntile_cap=128" "--tune umma_ntile_cap=64" "--tune umma_ntile_cap=32" "--tune umma_ntile_cap=96"; do
  echo "tune: $t"
  python tools/time_forward.py --batch 1 --frames 379 --reps 20 --parts 0 --graph 1 --pdl 0 $t 2>&1 | grep "ms/step" | awk 'NR%2==0'
done
python tools/time_forward.py --batch 1 --frames 379 --reps 20 --parts 0 --graph 1 --pdl 1 --tune umma_ntile_cap=64 2>&1 | grep "ms/step" | awk 'NR%2==0'
