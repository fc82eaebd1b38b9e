#!/bin/bash
mkdir -p gpurun_out
for v in 0 4; do
  echo "variant $v"
  DPFT_CONV_STREAM_VARIANT=$v timeout 100 python tools/conv_bench.py s2_conv2 s3_conv1 s3_conv2 s4_conv1 s4_conv2 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try:
        r=json.loads(l); print(r['layer'], r['us_cold'])
    except Exception: print(l[:200])"
  DPFT_CONV_STREAM_VARIANT=$v timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys,json
r=json.loads(sys.stdin.read()); print('ms', r['ms_per_step'], 'fps', r['value'], 'e2e', r['e2e']['value'], 'seq', r['sequential']['ms_per_step'])"
done
( DPFT_CONV_STREAM_VARIANT=4 timeout 200 python -m pytest tests/test_conv_gpu.py -m gpu -q -x --timeout 60 2>&1 | tail -2 )
