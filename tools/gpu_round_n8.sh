#!/bin/bash
set -u
OUT=gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/r02_bench_n${N}_a.json 2>$OUT/bench_n${N}_a.err; echo "bench n$N rc=$?"
tail -3 $OUT/bench_n${N}_a.err
python -c "
import json
l=json.loads(open('$OUT/r02_bench_n${N}_a.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','per_rank_ms','bf16_b128','hour','v2','clocks'):
    print(k, l.get(k))
"
