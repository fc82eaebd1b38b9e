#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/conv_debug.py > gpurun_out/conv_debug.jsonl 2> gpurun_out/conv_debug.err
timeout 600 python -m pytest tests/test_conv_gpu.py -q --timeout 120 -x 2>&1 | tail -5 > gpurun_out/pytest_conv.log
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_features_gpu.py -q --timeout 200 2>&1 | grep -E "AssertionError|passed|failed|FAILED|Error" > gpurun_out/pytest_model.log
timeout 600 python tools/stage_times.py > gpurun_out/stage_times.json 2> gpurun_out/stage_times.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
cut -c1-330 gpurun_out/conv_debug.jsonl; cat gpurun_out/pytest_conv.log gpurun_out/pytest_model.log; cat gpurun_out/stage_times.json; tail -3 gpurun_out/stage_times.err; cut -c1-400 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
