#!/bin/bash
mkdir -p gpurun_out
( timeout 400 python -m pytest tests/test_features_gpu.py tests/test_model_gpu.py tests/test_train_backbone_gpu.py -m gpu -q -x --timeout 120 2>&1 | tail -4 )
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys,json
r=json.loads(sys.stdin.read()); print('ms', r['ms_per_step'], 'fps', r['value'], 'e2e', r['e2e']['value'], 'seq', r['sequential']['ms_per_step'])"
timeout 200 python tools/stage_times.py 2>/dev/null | python -c "
import sys,json
r=json.loads(sys.stdin.read()); print({k:v for k,v in r.items() if 'stem' in k or 'forward' in k})"
