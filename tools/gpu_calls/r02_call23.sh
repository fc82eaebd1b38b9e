#!/bin/bash
O=gpurun_out/r02c23; mkdir -p $O
timeout 300 python -m pytest tests/test_model_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider -k "oracle_path" 2>&1 | grep -v "^E    *+ \|where " | tail -40 | cut -c1-250
DPFT_TRAIN_PARALLEL_VIEWS=0 timeout 300 python -m pytest tests/test_model_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider -k "oracle_path" 2>&1 | tail -3
