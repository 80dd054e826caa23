#!/bin/bash
# final-build numbers: GPU suite, bench line (N=1), per-op tables, smoke
set -u
OUT=gpurun_out
python -m pytest tests -m gpu -q > $OUT/r02_pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/r02_pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/r02_bench_reference.json 2>/dev/null; echo "reference rc=$?"
python bench.py --steps 20 --warmup 3 > $OUT/r02_bench_n1.json 2>$OUT/bench_n1.err; echo "bench rc=$?"; tail -2 $OUT/bench_n1.err
python tools/profile_ops.py --precision fp32 --out $OUT/r02_ops_fp32.txt > /dev/null 2>&1
python tools/profile_ops.py --precision bf16 --out $OUT/r02_ops_bf16.txt > /dev/null 2>&1
python tools/profile_diffsvc.py --brief 1 > $OUT/r02_diffsvc_ops.txt 2>&1; python tools/profile_diffsvc.py --batch 16 --frames 938 --brief 1 >> $OUT/r02_diffsvc_ops.txt 2>&1; cat $OUT/r02_diffsvc_ops.txt | grep -v "^  L"
for P in fp32 bf16; do
  python tools/ncu_diffsvc_target.py $P > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $OUT/launches_diffsvc_${P}.csv python tools/ncu_diffsvc_target.py $P > $OUT/ncu_diffsvc_${P}.log 2>&1; echo "diffsvc launch list $P rc=$?"
done
python tools/time_forward.py --v2 --batch 8 --frames 2584 --parts 2 2>&1 | grep -v Broken | head -4
python -c "
import json
l=json.loads(open('$OUT/r02_bench_n1.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','parity','vs_eager','diffsvc_step','class_ms_per_step','clocks','single_utterance_latency','gpu_launches'):
    print(k, l.get(k))
print('bf16', {k: l['bf16'][k] for k in ('value_per_gpu','ms_per_step','conv_frac','amp_frac')})
print('roofline', l['roofline']['frac'], l['roofline']['issued_frac'], l['roofline_amp']['frac'])
print('cpu', l['cpu_baseline']['value'], l['cpu_baseline']['cores'])
print('eager', {k: v for k, v in l['torch_eager_gpu'].items() if k != 'what'})
"
