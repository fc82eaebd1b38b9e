#!/bin/bash
mkdir -p gpurun_out
for f in "--side-priority" ""; do
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline $f > gpurun_out/bench_prio.json 2> gpurun_out/bench_prio.err
tail -1 gpurun_out/bench_prio.json | python -c "
import sys,json
r=json.loads(sys.stdin.read()); print('$f', 'ms', r['ms_per_step'], 'fps', r['value'], 'e2e', r['e2e']['value'], 'seq', r['sequential']['ms_per_step'])"
tail -2 gpurun_out/bench_prio.err
done
