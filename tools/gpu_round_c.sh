#!/bin/bash
set -u
OUT=gpurun_out
python -m pytest tests -m gpu -q > $OUT/r02_pytest_gpu_d.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/r02_pytest_gpu_d.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/r02_bench_n2_a.json 2>$OUT/bench_n2_a.err; echo "bench n2 rc=$?"
tail -5 $OUT/bench_n2_a.err
python -c "
import json
l=json.loads(open('$OUT/r02_bench_n2_a.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','per_rank_ms','bf16_b128','hour','v2','parity'):
    print(k, l.get(k))
"
