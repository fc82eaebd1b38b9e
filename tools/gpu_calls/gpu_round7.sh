#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_features_gpu.py tests/test_model_gpu.py -q --timeout 200 2>&1 | grep -E "AssertionError|passed|failed|FAILED|Error|rror" | head -20 > gpurun_out/pytest.log
timeout 600 python tools/stage_times.py > gpurun_out/stage_times.json 2> gpurun_out/stage_times.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'conv_gemm|decoder_|fpn_|stem_|maxpool|msda' -s 500 -c 260 --csv --log-file gpurun_out/launches_native.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
cat gpurun_out/pytest.log; cat gpurun_out/stage_times.json; tail -3 gpurun_out/stage_times.err; cut -c1-300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
