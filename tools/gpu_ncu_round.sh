#!/bin/bash
# One gpurun call: plain run, ncu launch list (time + DRAM bytes per launch) of one forward of the
# bench shape for both precisions, and --set full captures of three kernels of the fp32 path
# (wide conv, narrow conv, Activation1d).  Usage: tools/gpu_ncu_round.sh <tag>
set -u
TAG=${1:-v18}
OUT=gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for P in fp32 bf16; do
  python tools/ncu_target.py $P > $OUT/plain_${P}_$TAG.log 2>&1 || { echo "plain run failed ($P)"; tail -20 $OUT/plain_${P}_$TAG.log; exit 1; }
  # launches 0..(n-1) are the warm-up forward (+ weight packing); the second forward is the last 235 launches
  ncu --metrics $M --clock-control none --csv --log-file $OUT/launches_${P}_$TAG.csv python tools/ncu_target.py $P > $OUT/ncu_l_${P}_$TAG.log 2>&1
  echo "launch list $P rc=$?"
done
# second forward starts at conv_umma launch 115 / amp launch 109
ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 129 -c 1 -o $OUT/prof_conv_s0k11_$TAG -f python tools/ncu_target.py fp32 > $OUT/ncu_a_$TAG.log 2>&1
echo "wide conv rc=$?"
ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 224 -c 1 -o $OUT/prof_conv_s5k11_$TAG -f python tools/ncu_target.py fp32 > $OUT/ncu_b_$TAG.log 2>&1
echo "narrow conv rc=$?"
ncu --set full --clock-control none --import-source on -k regex:amp_kernel -s 127 -c 1 -o $OUT/prof_amp_s1_$TAG -f python tools/ncu_target.py fp32 > $OUT/ncu_c_$TAG.log 2>&1
echo "amp rc=$?"
ncu --set full --clock-control none --import-source on -k regex:amp_kernel -s 199 -c 1 -o $OUT/prof_amp_s5_$TAG -f python tools/ncu_target.py fp32 > $OUT/ncu_d_$TAG.log 2>&1
echo "amp C=24 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:amp_mma -s 127 -c 1 -o $OUT/prof_ampmma_s1_bf16_$TAG -f python tools/ncu_target.py bf16 > $OUT/ncu_e_$TAG.log 2>&1
echo "amp_mma bf16 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 148 -c 1 -o $OUT/prof_conv_s1k11_$TAG -f python tools/ncu_target.py fp32 > $OUT/ncu_f_$TAG.log 2>&1
echo "stage-1 conv rc=$?"
ls -la $OUT | tail -12
