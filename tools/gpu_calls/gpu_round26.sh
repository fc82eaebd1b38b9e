#!/bin/bash
# two GPUs: replica inference bench and the graphed data-parallel training step (NCCL all-reduce captured in the graph)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_infer_2gpu.json 2> gpurun_out/bench_infer_2gpu.err
tail -1 gpurun_out/bench_infer_2gpu.json | cut -c1-300; tail -3 gpurun_out/bench_infer_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --mode train --steps 10 --warmup 3 > gpurun_out/bench_train_dp2.json 2> gpurun_out/bench_train_dp2.err
tail -1 gpurun_out/bench_train_dp2.json; tail -5 gpurun_out/bench_train_dp2.err
