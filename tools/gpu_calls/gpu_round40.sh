#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py tests/test_train_backbone_gpu.py -m gpu -q -x --timeout 120 2>&1 | tail -3 )
timeout 100 python tools/conv_bench.py s1_conv2 2>&1 | cut -c1-220
for v in 1 0; do
DPFT_CONV_HALO=$v timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_halo$v.json 2> gpurun_out/bench_halo$v.err
tail -1 gpurun_out/bench_halo$v.json | python -c "
import sys,json
r=json.loads(sys.stdin.read()); print('halo=$v', 'ms', r['ms_per_step'], 'fps', r['value'], 'e2e', r['e2e']['value'], 'seq', r['sequential']['ms_per_step'], 'roof', r['roofline']['frac'])"
done
timeout 600 python bench.py --mode train --steps 10 --warmup 3 2>/dev/null | tail -1 | cut -c1-160
