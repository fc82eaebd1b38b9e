#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_features_gpu.py tests/test_model_gpu.py tests/test_train_backbone_gpu.py tests/test_train_ops_gpu.py -m gpu -x -q --timeout 200 ) > gpurun_out/pytest_feat.log 2>&1
tail -5 gpurun_out/pytest_feat.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_epi8.json 2> gpurun_out/bench_epi8.err
tail -1 gpurun_out/bench_epi8.json | python -c "
import sys,json
r=json.loads(sys.stdin.read()); print('ms', r['ms_per_step'], 'fps', r['value'], 'e2e', r['e2e']['value'], 'seq', r['sequential']['ms_per_step'], 'roof', r['roofline']['frac'])"
timeout 600 python tools/stage_times.py 2>/dev/null | python -c "
import sys,json
r=json.loads(sys.stdin.read()); print({k:v for k,v in r.items() if 'camera' in k or 'forward' in k or 'decoder' in k})"
