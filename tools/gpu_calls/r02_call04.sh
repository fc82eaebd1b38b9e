#!/bin/bash
# Round 2, call 4: what bounds the expand layers?  Weight-stationary variants (tile width, A stages, ring depth, epilogue warps)
# and the same GEMM without the residual stream.
O=gpurun_out/r02c04; mkdir -p $O
for v in 0 1 2 3 4 5; do
DPFT_WS_VARIANT=$v timeout 120 python tools/conv_bench.py --no-lib s2_conv3 s3_conv3 2>&1 | cut -c1-200 | sed "s/^/wsv=$v /"
done | tee $O/ws_variants.txt
DPFT_CONV_WS=0 timeout 120 python tools/conv_bench.py --no-lib s3_conv3 s3_conv3_nores 2>&1 | cut -c1-200 | sed "s/^/generic /" | tee -a $O/ws_variants.txt
for v in 1 2 3; do
DPFT_CONV_WS=0 DPFT_CONV_STREAM_VARIANT=$v timeout 120 python tools/conv_bench.py --no-lib s3_conv3 2>&1 | cut -c1-200 | sed "s/^/generic stream variant $v /"
done | tee -a $O/ws_variants.txt
