#!/bin/bash
set -u
timeout 300 python -m pytest tests/test_gpu_generator.py -m gpu -q -k "synthesis or pcm16" 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 --no-eager --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value', l['value'], 'ms', l['ms_per_step'], 'e2e', l['e2e']['value'], 'lat', l['single_utterance_latency'])
"
