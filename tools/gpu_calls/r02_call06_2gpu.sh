#!/bin/bash
# Round 2, call 6 (2 GPUs): NCCL gradient-bucket test, the default bench line at N=2 as the driver launches it (inference
# replicas + the config-4 training step with and without its all-reduces), wall time of the whole bench.
O=gpurun_out/r02c06; mkdir -p $O
nvidia-smi -L
timeout 300 python -m pytest tests/test_ddp_nccl_gpu.py -m gpu -q --timeout 200 -p no:cacheprovider 2>&1 | tail -5
T0=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.out 2> $O/bench_n2.err
echo "bench N=2 rc=$? wall=$(( $(date +%s) - T0 )) s"
tail -1 $O/bench_n2.out > $O/bench_n2.json; tail -5 $O/bench_n2.err
python - <<'PY'
import json
r = json.load(open('gpurun_out/r02c06/bench_n2.json'))
print({k: r[k] for k in ('value', 'ms_per_step', 'n_gpus')}, 'e2e', r['e2e']['value'])
print('train', json.dumps(r['train'])[:900])
PY
T0=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
echo "reference arm N=2 wall=$(( $(date +%s) - T0 )) s"
