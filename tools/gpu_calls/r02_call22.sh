#!/bin/bash
# Round 2, call 22: weight gradients on their own stream: training tests, step A/B.
O=gpurun_out/r02c22; mkdir -p $O
timeout 600 python -m pytest tests/test_train_step.py tests/test_train_backbone_gpu.py tests/test_golden_taps_gpu.py tests/test_model_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -5 | cut -c1-300
for v in 0 1 0 1; do
DPFT_WGRAD_STREAM=$v timeout 300 python bench.py --mode train --steps 20 --warmup 3 2>/dev/null | tail -1 | python -c "
import sys, json
r = json.loads(sys.stdin.read())
print('wgrad_stream=$v ms', round(r['ms_per_step'], 3), 'fps', round(r['value'], 1), 'loss', r['final_loss'])"
done | tee $O/wgrad_stream_ab.txt
