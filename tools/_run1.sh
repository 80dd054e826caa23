python tools/ncu_target.py fp32 > gpurun_out/plain_a.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 224 -c 1 -o gpurun_out/prof_conv_s5k11_v6 -f python tools/ncu_target.py fp32 > gpurun_out/ncu_b.log 2>&1
timeout 300 python -m pytest tests -m gpu -q -k "amp_packed" 2>&1 | tail -3
