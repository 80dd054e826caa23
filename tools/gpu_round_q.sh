#!/bin/bash
set -u
BVG_B200_LIB=$PWD/svc_inference_pipeline_b200/libbvg_b200_epi16.so timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "conv" 2>&1 | tail -2
for rep in 1 2; do
  for lib in libbvg_b200.so libbvg_b200_epi16.so; do
    echo "== $lib"
    BVG_B200_LIB=$PWD/svc_inference_pipeline_b200/$lib timeout 300 python tools/profile_ops.py --precision fp32 | grep -E "#  conv  L(60032|120064|240128)|total="
  done
done
for lib in libbvg_b200.so libbvg_b200_epi16.so; do
  echo "== bf16 $lib"
  BVG_B200_LIB=$PWD/svc_inference_pipeline_b200/$lib timeout 300 python tools/profile_ops.py --precision bf16 | grep -E "#  conv  L(60032|120064|240128)|total="
done
