#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/accuracy_report.py > gpurun_out/accuracy.jsonl 2> gpurun_out/accuracy.err
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_features_gpu.py -q --timeout 200 2>&1 | grep -E "AssertionError|passed|failed|FAILED|Error" > gpurun_out/pytest_model.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 600 python tools/stage_times.py > gpurun_out/stage_times.json 2> gpurun_out/stage_times.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'conv_gemm|decoder_|fpn_|stem_|maxpool|msda' -s 500 -c 260 --csv --log-file gpurun_out/launches_native.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
cut -c1-2000 gpurun_out/accuracy.jsonl | head -3; cat gpurun_out/pytest_model.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/stage_times.json; tail -3 gpurun_out/stage_times.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
