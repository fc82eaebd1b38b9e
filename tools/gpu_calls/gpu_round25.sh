#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_attention_gpu.py tests/test_conv_gpu.py -m gpu -x -q --timeout 120 ) > gpurun_out/pytest_attention.log 2>&1
tail -5 gpurun_out/pytest_attention.log
timeout 300 python tools/attention_bench.py > gpurun_out/attention_bench.jsonl 2> gpurun_out/attention_bench.err
python - <<'PY'
import json
for l in open('gpurun_out/attention_bench.jsonl'):
    r=json.loads(l); print(r['case'], {k:round(v,1) for k,v in r.items() if k.endswith('_us')}, 'err', r['err_precise'])
PY
tail -3 gpurun_out/attention_bench.err
DPFT_CONV_PDL=0 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_pdl0.json 2> gpurun_out/bench_pdl0.err
DPFT_CONV_PDL=1 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_pdl1.json 2> gpurun_out/bench_pdl1.err
python - <<'PY'
import json
for n in ('bench_pdl0','bench_pdl1'):
    try:
        r=json.loads(open(f'gpurun_out/{n}.json').read().strip().splitlines()[-1]); print(n, r['ms_per_step'], r['value'], r['e2e']['value'], r['roofline']['frac'], r['roofline']['ms_in_kernel_per_step'])
    except Exception as e: print(n, 'ERR', e)
PY
tail -3 gpurun_out/bench_pdl1.err
