#!/bin/bash
# Round-2 ncu evidence (one gpurun call): plain run, launch list (time + DRAM bytes per launch) of one forward of the
# bench shape for both precisions, and --set full captures of the hot kernels.  Usage: tools/gpu_ncu_round2.sh <tag>
# Launch order per forward (fp32 path): 59 conv_pair_kernel launches (conv_pre, ups.0, 18 x C=768, ups.1, 18 x C=384,
# ups.2, 18 x C=192, ups.3), 56 conv_umma_kernel launches (18 x C=96, ups.4, 18 x C=48, ups.5, 18 x C=24), 109 amp.
set -u
TAG=${1:-r02}
OUT=gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for P in fp32 bf16; do
  python tools/ncu_target.py $P > $OUT/plain_${P}_$TAG.log 2>&1 || { echo "plain run failed ($P)"; tail -20 $OUT/plain_${P}_$TAG.log; exit 1; }
  ncu --metrics $M --clock-control none --csv --log-file $OUT/launches_${P}_$TAG.csv python tools/ncu_target.py $P > $OUT/ncu_l_${P}_$TAG.log 2>&1
  echo "launch list $P rc=$?"
done
cap() {  # name, kernel regex, skip, precision
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o $OUT/prof_$1_$TAG -f python tools/ncu_target.py $4 > $OUT/ncu_$1_$TAG.log 2>&1
  echo "$1 rc=$?"
}
cap pair_s0k11_fp32 conv_pair 73 fp32     # resblocks.2.convs1.0  768->768 k11
cap pair_s1k11_fp32 conv_pair 92 fp32     # resblocks.5.convs1.0  384->384 k11
cap pair_s2k11_fp32 conv_pair 111 fp32    # resblocks.8.convs1.0  192->192 k11
cap umma_s5k11_fp32 conv_umma 106 fp32    # resblocks.17.convs1.0 24->24 k11 (time-folded)
cap amp_s1_fp32 amp_kernel 127 fp32       # stage-1 Activation1d C=384
cap pair_s1k11_bf16 conv_pair 92 bf16
cap pair_s2k11_bf16 conv_pair 111 bf16
cap ampmma_s1_bf16 amp_mma 127 bf16
ls -la $OUT | tail -12
