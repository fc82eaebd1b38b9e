#!/bin/bash
mkdir -p gpurun_out
for bo in 1 0; do
echo "DPFT_HALO_BASE_OFFSET=$bo"
( DPFT_HALO_BASE_OFFSET=$bo timeout 120 python -m pytest tests/test_conv_gpu.py -m gpu -q --timeout 60 -k "halo" 2>&1 | tail -4 )
DPFT_HALO_BASE_OFFSET=$bo timeout 100 python tools/conv_bench.py s1_conv2 2>&1 | cut -c1-220
done
