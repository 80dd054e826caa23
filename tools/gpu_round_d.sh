#!/bin/bash
set -u
OUT=gpurun_out
python -m pytest tests/test_gpu_ops.py tests/test_gpu_generator.py -m gpu -q -x > $OUT/r02_pytest_gpu_e.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/r02_pytest_gpu_e.log
for rep in 1 2; do
  for lib in libbvg_b200.so libbvg_b200_nopair.so; do
    echo "== $lib"
    BVG_B200_LIB=$PWD/svc_inference_pipeline_b200/$lib python tools/profile_ops.py --precision bf16 | grep "#  amp"
    BVG_B200_LIB=$PWD/svc_inference_pipeline_b200/$lib python tools/time_forward.py --precisions bf16 --parts 2 | head -1
  done
done
