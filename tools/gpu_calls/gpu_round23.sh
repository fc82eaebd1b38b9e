#!/bin/bash
# graphed training step (tests + bench), ncu launch list of one eval step, ncu DRAM traffic / tensor-pipe activity of every conv launch
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_train_step.py -m gpu -x -q ) > gpurun_out/pytest_train_step.log 2>&1
tail -15 gpurun_out/pytest_train_step.log
timeout 600 python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/bench_train_graph.json 2> gpurun_out/bench_train_graph.err
cat gpurun_out/bench_train_graph.json; tail -5 gpurun_out/bench_train_graph.err
# second forward only (-s = launches of the first forward incl. one-off packing kernels are skipped by counting from the end in post-processing)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_eval_step.csv python tools/one_forward.py 2 > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum --clock-control none -k regex:conv_gemm -c 500 --csv --log-file gpurun_out/conv_traffic.csv python tools/one_forward.py 2 > gpurun_out/ncu_conv_traffic.log 2>&1
tail -2 gpurun_out/ncu_conv_traffic.log
wc -l gpurun_out/launches_eval_step.csv gpurun_out/conv_traffic.csv
