#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 4 -c 1 -f -o gpurun_out/conv_s3c3 python tools/conv_bench.py s3_conv3 > gpurun_out/ncu_s3c3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 4 -c 1 -f -o gpurun_out/conv_s3c1 python tools/conv_bench.py s3_conv1 > gpurun_out/ncu_s3c1.log 2>&1
tail -3 gpurun_out/ncu_s3c3.log gpurun_out/ncu_s3c1.log; ls -la gpurun_out/*.ncu-rep
