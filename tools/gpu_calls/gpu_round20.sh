#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/stage_times.py > gpurun_out/stage_times.json 2> gpurun_out/stage_times.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 python tools/accuracy_report.py > gpurun_out/accuracy.jsonl 2> gpurun_out/accuracy.err
cat gpurun_out/stage_times.json; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err; python - <<'PY'
import json
for l in open('gpurun_out/accuracy.jsonl'):
    r=json.loads(l); print(r['case'], 'f16', r['f16'], 'bf16', r['bf16'])
PY
tail -2 gpurun_out/accuracy.err
