#!/bin/bash
# Round 2, call 17: training forward/backward of the extra views on forked streams: parity tests, then A/B of the step.
O=gpurun_out/r02c17; mkdir -p $O
timeout 600 python -m pytest tests/test_train_step.py tests/test_train_backbone_gpu.py tests/test_train_ops_gpu.py tests/test_model_gpu.py tests/test_golden_taps_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider -x 2>&1 | tail -4
for v in 0 1 0 1; do
DPFT_TRAIN_PARALLEL_VIEWS=$v timeout 300 python bench.py --mode train --steps 20 --warmup 3 2>/dev/null | tail -1 | python -c "
import sys, json
r = json.loads(sys.stdin.read())
print('train_parallel_views=$v ms', round(r['ms_per_step'], 3), 'fps', round(r['value'], 1), 'loss', r['final_loss'])"
done | tee $O/train_parallel_views_ab.txt
