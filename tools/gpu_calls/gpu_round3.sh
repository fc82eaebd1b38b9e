#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/accuracy_report.py > gpurun_out/accuracy.jsonl 2> gpurun_out/accuracy.err
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_features_gpu.py -q --timeout 120 2>&1 | tail -30 > gpurun_out/pytest_conv_features.log
timeout 900 python -m pytest tests/test_model_gpu.py -q --timeout 200 2>&1 | grep -E "AssertionError|passed|failed|FAILED" > gpurun_out/pytest_model.log
timeout 300 python tools/msda_sweep.py > gpurun_out/msda_sweep.jsonl 2> gpurun_out/msda_sweep.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/launches_native.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
cat gpurun_out/accuracy.jsonl; tail -5 gpurun_out/accuracy.err; tail -12 gpurun_out/pytest_conv_features.log; cat gpurun_out/pytest_model.log; cat gpurun_out/msda_sweep.jsonl | cut -c1-330; cat gpurun_out/bench.json
