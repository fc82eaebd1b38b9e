#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:fpn_output_tc_kernel -s 2 -c 1 -f -o gpurun_out/fpn_out_cam python tools/one_forward.py 2 > gpurun_out/ncu_fpn.log 2>&1
tail -n 2 gpurun_out/ncu_fpn.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_tile.json 2> gpurun_out/bench_tile.err
tail -1 gpurun_out/bench_tile.json | python -c "
import sys,json
r=json.loads(sys.stdin.read()); print('ms', r['ms_per_step'], 'fps', r['value'], 'e2e', r['e2e']['value'], 'seq', r['sequential']['ms_per_step'], 'roof', r['roofline']['frac'])"
