#!/bin/bash
mkdir -p gpurun_out
( DPFT_CONV_STREAM_VARIANT=4 timeout 600 python -m pytest tests/test_conv_gpu.py -m gpu -x -q --timeout 120 ) > gpurun_out/pytest_conv_eg4.log 2>&1
tail -4 gpurun_out/pytest_conv_eg4.log
for v in 0 3 4; do
  echo "variant $v"
  DPFT_CONV_STREAM_VARIANT=$v timeout 200 python tools/conv_bench.py s1_conv1 s1_conv2 s1_conv3 s2_conv3 s3_conv2 s3_conv3 s4_conv3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try:
        r=json.loads(l); print(r['layer'], r['us_cold'], r['tflops_cold'], r['gbps_cold'])
    except Exception: print(l[:200])"
done > gpurun_out/conv_eg4_ab.txt
cat gpurun_out/conv_eg4_ab.txt
for v in 0 3 4; do
DPFT_CONV_STREAM_VARIANT=$v timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err
done
python - <<'PY'
import json
for n in ('bench_v0','bench_v3','bench_v4'):
    try:
        r=json.loads(open(f'gpurun_out/{n}.json').read().strip().splitlines()[-1]); print(n, 'ms', r['ms_per_step'], 'fps', r['value'], 'e2e', r['e2e']['value'], 'seq ms', r['sequential']['ms_per_step'], 'roof', r['roofline']['frac'], r['roofline']['ms_in_kernel_per_step'])
    except Exception as e: print(n, 'ERR', e)
PY
