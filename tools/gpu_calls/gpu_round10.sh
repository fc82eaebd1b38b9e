#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py -q --timeout 120 -x 2>&1 | tail -3 > gpurun_out/pytest_conv.log
timeout 600 python -m pytest tests/test_msda_gpu.py tests/test_features_gpu.py tests/test_model_gpu.py -q --timeout 200 2>&1 | grep -E "AssertionError|passed|failed|FAILED|Error|rror" | head -20 > gpurun_out/pytest.log
timeout 300 python tools/msda_sweep.py > gpurun_out/msda_sweep.jsonl 2> gpurun_out/msda_sweep.err
timeout 300 python tools/conv_bench.py > gpurun_out/conv_bench.jsonl 2> gpurun_out/conv_bench.err
timeout 600 python tools/stage_times.py > gpurun_out/stage_times.json 2> gpurun_out/stage_times.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/pytest_conv.log gpurun_out/pytest.log; python - <<'PY'
import json
for l in open('gpurun_out/msda_sweep.jsonl'):
    r=json.loads(l); print(r['case'], 'fwd %.1fus %.3f'%(r['fwd_us'],r['fwd_frac']), 'bwd %.1fus %.3f'%(r['bwd_us'],r['bwd_frac']))
for l in open('gpurun_out/conv_bench.jsonl'):
    r=json.loads(l); print(r['layer'], r['us_cold'], 'us', r['tflops_cold'], 'TF', r['gbps_cold'], 'GB/s')
PY
tail -3 gpurun_out/conv_bench.err; cat gpurun_out/stage_times.json; cut -c1-250 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
