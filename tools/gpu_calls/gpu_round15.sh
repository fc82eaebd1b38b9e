#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus2.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/infer2.json 2> gpurun_out/infer2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --mode train --steps 5 --warmup 3 > gpurun_out/train2.json 2> gpurun_out/train2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/ref2.json 2> gpurun_out/ref2.err
cat gpurun_out/gpus2.txt; cut -c1-330 gpurun_out/infer2.json; tail -3 gpurun_out/infer2.err; cat gpurun_out/train2.json; tail -3 gpurun_out/train2.err; cut -c1-300 gpurun_out/ref2.json; tail -2 gpurun_out/ref2.err
