#!/bin/bash
# One gpurun call: plain run, ncu launch list of one forward, and --set full captures of three kernels
# (wide conv, narrow conv, AMP) of the fp32 bench shape.  Usage: tools/gpu_ncu_round.sh <tag>
set -u
TAG=${1:-v5}
OUT=gpurun_out
python tools/ncu_target.py fp32 > $OUT/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -20 $OUT/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 227 -c 240 --csv --log-file $OUT/launches_fp32_$TAG.csv python tools/ncu_target.py fp32 > $OUT/ncu_l_$TAG.log 2>&1
echo "launch list rc=$?"
# second forward starts at conv_umma launch 115 / amp launch 109
ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 129 -c 1 -o $OUT/prof_conv_s0k11_$TAG -f python tools/ncu_target.py fp32 > $OUT/ncu_a_$TAG.log 2>&1
echo "wide conv rc=$?"
ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 224 -c 1 -o $OUT/prof_conv_s5k11_$TAG -f python tools/ncu_target.py fp32 > $OUT/ncu_b_$TAG.log 2>&1
echo "narrow conv rc=$?"
ncu --set full --clock-control none --import-source on -k regex:amp_kernel -s 127 -c 1 -o $OUT/prof_amp_s1_$TAG -f python tools/ncu_target.py fp32 > $OUT/ncu_c_$TAG.log 2>&1
echo "amp rc=$?"
ls -la $OUT
