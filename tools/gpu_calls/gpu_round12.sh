#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/conv_bench.py s1_conv3 s3_conv3 s3_conv1 > gpurun_out/conv_bench.jsonl 2> gpurun_out/conv_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 4 -c 1 -o gpurun_out/conv_s1c3 python tools/conv_bench.py s1_conv3 > gpurun_out/ncu_conv.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:msda_fwd -s 3 -c 1 -o gpurun_out/msda_ring python tools/msda_sweep.py --only cfg5_bf16_D32 --reps 2 >> gpurun_out/ncu_conv.log 2>&1
cat gpurun_out/conv_bench.jsonl
