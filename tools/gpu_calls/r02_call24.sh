#!/bin/bash
O=gpurun_out/r02c24; mkdir -p $O
for rep in 1 2 3 4; do
timeout 400 python -m pytest tests/test_train_step.py tests/test_train_backbone_gpu.py tests/test_golden_taps_gpu.py tests/test_model_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider > $O/rep$rep.txt 2>&1
tail -2 $O/rep$rep.txt
grep -n "^E  .*assert\|AssertionError\|^FAILED" $O/rep$rep.txt | grep -v "where" | cut -c1-300 | head -8
done
