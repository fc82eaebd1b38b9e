#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/conv_bench.py --sweep > gpurun_out/conv_sweep.jsonl 2> gpurun_out/conv_sweep.err
cat gpurun_out/conv_sweep.jsonl; tail -3 gpurun_out/conv_sweep.err
