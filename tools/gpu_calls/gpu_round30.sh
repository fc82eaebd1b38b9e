#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py tests/test_features_gpu.py -m gpu -x -q --timeout 120 ) > gpurun_out/pytest_conv.log 2>&1
tail -6 gpurun_out/pytest_conv.log
for v in 0 1; do
  echo "DPFT_CONV_EXPAND=$v"
  DPFT_CONV_EXPAND=$v timeout 200 python tools/conv_bench.py s1_conv3 s2_conv3 s3_conv3 2>&1 | cut -c1-200
done > gpurun_out/conv_expand_ab.txt
cat gpurun_out/conv_expand_ab.txt
DPFT_CONV_EXPAND=0 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_expand0.json 2> gpurun_out/bench_expand0.err
DPFT_CONV_EXPAND=1 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_expand1.json 2> gpurun_out/bench_expand1.err
python - <<'PY'
import json
for n in ('bench_expand0','bench_expand1'):
    try:
        r=json.loads(open(f'gpurun_out/{n}.json').read().strip().splitlines()[-1]); print(n, r['ms_per_step'], r['value'], r['e2e']['value'], r['roofline']['frac'], r['roofline']['ms_in_kernel_per_step'])
    except Exception as e: print(n, 'ERR', e)
PY
tail -3 gpurun_out/bench_expand1.err
