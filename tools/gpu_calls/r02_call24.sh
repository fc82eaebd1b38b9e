#!/bin/bash
O=gpurun_out/r02c24; mkdir -p $O
for rep in 1 2 3; do
timeout 400 python -m pytest tests/test_train_step.py tests/test_train_backbone_gpu.py tests/test_golden_taps_gpu.py tests/test_model_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider > $O/rep$rep.txt 2>&1
tail -2 $O/rep$rep.txt
grep -n "^E  .*assert\|AssertionError\|^FAILED" $O/rep$rep.txt | grep -v "where" | cut -c1-300 | head -8
done
for v in 0 1; do
DPFT_TRAIN_PARALLEL_VIEWS=$v DPFT_WGRAD_STREAM=$v timeout 300 python bench.py --mode train --steps 20 --warmup 3 2>/dev/null | tail -1 | python -c "
import sys, json
r = json.loads(sys.stdin.read())
print('forked+wgrad_stream=$v ms', round(r['ms_per_step'], 3), 'fps', round(r['value'], 1), 'loss', r['final_loss'])"
done
