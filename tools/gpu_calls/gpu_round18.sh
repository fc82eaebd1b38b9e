#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 4 -c 1 -o gpurun_out/conv_s1c3_v2 python tools/conv_bench.py s1_conv3 > gpurun_out/ncu_conv.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:decoder_layer -s 2 -c 1 -o gpurun_out/decoder_layer python bench.py --steps 1 --warmup 1 --no-cpu-baseline >> gpurun_out/ncu_conv.log 2>&1
tail -3 gpurun_out/ncu_conv.log
