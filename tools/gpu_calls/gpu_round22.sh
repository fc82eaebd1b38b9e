#!/bin/bash
# re-entry check: full GPU test suite, headline bench, native training bench, training kernel breakdown
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --mode train --steps 5 --warmup 3 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err
cat gpurun_out/bench_train.json; tail -3 gpurun_out/bench_train.err
timeout 300 python tools/train_times.py --full > gpurun_out/train_times_full.txt 2> gpurun_out/train_times_full.err
cat gpurun_out/train_times_full.txt; tail -3 gpurun_out/train_times_full.err
