#!/bin/bash
# last call of the round: the streaming FPN output kernel — its own test first (short timeout), then the feature / model tests and an A/B
timeout 70 python -m pytest tests/test_features_gpu.py -m gpu -q -x --timeout 30 -k "streaming" 2>&1 | tail -3
if [ "${PIPESTATUS[0]}" != "0" ]; then echo "STREAMING TEST FAILED"; exit 0; fi
timeout 80 python -m pytest tests/test_features_gpu.py tests/test_model_gpu.py -m gpu -q -x --timeout 60 2>&1 | tail -2
for v in 1 0; do
DPFT_FPN_STREAM=$v timeout 70 python bench.py --steps 20 --warmup 4 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys,json
r=json.loads(sys.stdin.read()); print('fpn_stream=$v', 'ms', r['ms_per_step'], 'fps', r['value'], 'e2e', r['e2e']['value'], 'seq', r['sequential']['ms_per_step'])"
done
