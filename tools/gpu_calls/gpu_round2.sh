#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python tools/conv_debug.py > gpurun_out/conv_debug.jsonl 2> gpurun_out/conv_debug.err
timeout 600 python -m pytest tests/test_features_gpu.py -q --timeout 120 2>&1 | tail -40 > gpurun_out/pytest_features.log
timeout 900 python -m pytest tests/test_model_gpu.py -q --timeout 200 2>&1 | tail -60 > gpurun_out/pytest_model.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/conv_debug.jsonl; tail -15 gpurun_out/pytest_features.log; tail -25 gpurun_out/pytest_model.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
