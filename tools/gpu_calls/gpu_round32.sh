#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_infer_2gpu.json 2> gpurun_out/bench_infer_2gpu.err
echo "rc=$?"; tail -1 gpurun_out/bench_infer_2gpu.json | python -c "
import sys,json
r=json.loads(sys.stdin.read()); print(r['n_gpus'], r['ms_per_step'], r['value'], r['e2e']['value'], r['sequential'])"
tail -2 gpurun_out/bench_infer_2gpu.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --mode train --steps 8 --warmup 3 > gpurun_out/bench_train_dp2.json 2> gpurun_out/bench_train_dp2.err
echo "rc=$?"; tail -1 gpurun_out/bench_train_dp2.json | cut -c1-200
