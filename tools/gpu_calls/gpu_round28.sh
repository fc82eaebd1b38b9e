#!/bin/bash
mkdir -p gpurun_out
for v in 0 1 2; do
  echo "variant $v"
  DPFT_CONV_STREAM_VARIANT=$v timeout 200 python tools/conv_bench.py --sweep s1_conv3 s2_conv3 s3_conv3 s4_conv3 2>&1 | tail -4
done > gpurun_out/conv_stream_variants.txt
cat gpurun_out/conv_stream_variants.txt
