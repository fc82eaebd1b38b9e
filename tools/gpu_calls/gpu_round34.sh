#!/bin/bash
# end-of-round validation: full GPU suite, smoke, headline bench (with CPU baseline), training bench, stage times, launch list
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 ) > gpurun_out/pytest_gpu.log 2>&1
tail -6 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -1 gpurun_out/bench.json | python -c "
import sys,json
r=json.loads(sys.stdin.read()); print('ms', r['ms_per_step'], 'fps', r['value'], 'e2e', r['e2e']['value'], 'seq', r['sequential']['ms_per_step'], 'roof', r['roofline']['frac'], 'cpu', r.get('cpu_baseline',{}).get('value'), r['clocks'])"
tail -2 gpurun_out/bench.err
timeout 600 python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err
tail -1 gpurun_out/bench_train.json | cut -c1-160
timeout 300 python tools/stage_times.py > gpurun_out/stage_times.json 2> gpurun_out/stage_times.err; cat gpurun_out/stage_times.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_eval_step.csv python tools/one_forward.py 2 > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --profile-from-start off -k 'regex:conv_gemm|conv3x3_halo' --csv --log-file gpurun_out/conv_traffic.csv python tools/one_forward.py 1 > gpurun_out/ncu_conv_traffic.log 2>&1
wc -l gpurun_out/launches_eval_step.csv gpurun_out/conv_traffic.csv
