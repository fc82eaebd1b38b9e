#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --mode train --steps 5 --warmup 3 > gpurun_out/train1.json 2> gpurun_out/train1.err
cat gpurun_out/train1.json; tail -5 gpurun_out/train1.err
