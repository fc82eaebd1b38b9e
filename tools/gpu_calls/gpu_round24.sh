#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_attention_gpu.py -m gpu -x -q --timeout 120 ) > gpurun_out/pytest_attention.log 2>&1
tail -25 gpurun_out/pytest_attention.log
( timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -x -q ) > gpurun_out/pytest_model.log 2>&1
tail -5 gpurun_out/pytest_model.log
timeout 300 python tools/attention_bench.py > gpurun_out/attention_bench.jsonl 2> gpurun_out/attention_bench.err
cat gpurun_out/attention_bench.jsonl; tail -3 gpurun_out/attention_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_eval_step.csv python tools/one_forward.py 2 > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log; wc -l gpurun_out/launches_eval_step.csv
