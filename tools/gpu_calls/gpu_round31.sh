#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_infer_stream_gpu.py tests/test_model_gpu.py -m gpu -x -q --timeout 200 ) > gpurun_out/pytest_stream.log 2>&1
tail -12 gpurun_out/pytest_stream.log
for d in 2 3; do
timeout 600 python bench.py --steps 40 --warmup 6 --no-cpu-baseline --depth $d > gpurun_out/bench_depth$d.json 2> gpurun_out/bench_depth$d.err
tail -3 gpurun_out/bench_depth$d.err
done
python - <<'PY'
import json
for n in ('bench_depth2','bench_depth3'):
    try:
        r=json.loads(open(f'gpurun_out/{n}.json').read().strip().splitlines()[-1]); print(n, 'ms', r['ms_per_step'], 'fps', r['value'], 'e2e', r['e2e']['value'], 'seq', r['sequential'], r['clocks'])
    except Exception as e: print(n, 'ERR', e)
PY
