#!/bin/bash
# Round-2 GPU call A: GPU test-suite, default-build per-op tables (both paths), bench line.
set -u
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q -s > $OUT/r02_pytest_gpu_a.log 2>&1; echo "pytest rc=$?"
tail -5 $OUT/r02_pytest_gpu_a.log
grep -E "max-abs|SNR|recipe|large-argument" $OUT/r02_pytest_gpu_a.log | head -40
python tools/profile_ops.py --precision fp32 --out $OUT/r02_ops_fp32_a.txt > /dev/null 2>$OUT/ops_fp32.err; echo "ops fp32 rc=$?"
python tools/profile_ops.py --precision bf16 --out $OUT/r02_ops_bf16_a.txt > /dev/null 2>$OUT/ops_bf16.err; echo "ops bf16 rc=$?"
tail -22 $OUT/r02_ops_fp32_a.txt
python bench.py --steps 10 --warmup 3 > $OUT/r02_bench_a.json 2>$OUT/bench_a.err; echo "bench rc=$?"
tail -3 $OUT/bench_a.err
python -c "
import json
l=json.loads(open('$OUT/r02_bench_a.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','parity','torch_eager_gpu','vs_eager','bf16','class_ms_per_step','clocks'):
    print(k, l.get(k))
print('roofline', l['roofline']['frac'], l['roofline_amp']['frac'])
"
