#!/bin/bash
set -u
OUT=gpurun_out
python tools/conv_precision_diag.py 2>&1 | tail -5
python tools/precision_diag.py 2>&1 | tail -4
python -m pytest tests -m gpu -q -x > $OUT/r02_pytest_gpu_c.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/r02_pytest_gpu_c.log
python tools/profile_ops.py --precision fp32 --out $OUT/r02_ops_fp32_b.txt | tail -16
python tools/time_forward.py --precisions fp32 --parts 2
