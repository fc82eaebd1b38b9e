#!/bin/bash
mkdir -p gpurun_out
( timeout 120 python -m pytest tests/test_conv_gpu.py -m gpu -q --timeout 60 -k "halo" 2>&1 | tail -4 )
timeout 100 python tools/conv_bench.py s1_conv2 2>&1 | cut -c1-220
