#!/bin/bash
# Round 2, call 21 (2 GPUs): forked training streams under data parallelism: NCCL bucket test, training-step tests, N=2 bench.
O=gpurun_out/r02c21; mkdir -p $O
timeout 400 python -m pytest tests/test_ddp_nccl_gpu.py tests/test_train_step.py -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -6 | cut -c1-300
T0=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 20 --warmup 5 --train-timeout 200 > $O/bench_n2.out 2> $O/bench_n2.err
echo "bench N=2 rc=$? wall=$(( $(date +%s) - T0 )) s"
tail -1 $O/bench_n2.out > $O/bench_n2.json; grep -v "OMP_NUM_THREADS\|\*\*\*\*\|UserWarning\|scale = " $O/bench_n2.err | tail -4
python - <<'PY'
import json
r = json.load(open('gpurun_out/r02c21/bench_n2.json'))
print({k: r[k] for k in ('value', 'ms_per_step', 'n_gpus')}, 'e2e', r['e2e']['value'], r['e2e']['ms_per_step'])
t = r['train']
print('train', {k: t.get(k) for k in ('value', 'ms_per_step', 'allreduce_ms_exposed', 'allreduce_ms_exposed_raw', 'ms_per_step_without_allreduce', 'final_loss', 'error')})
PY
for v in 0 1; do
DPFT_TRAIN_PARALLEL_VIEWS=$v timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29580 + v)) bench.py --mode train --gpus 2 --steps 20 --warmup 3 2>/dev/null | tail -1 | python -c "
import sys, json
r = json.loads(sys.stdin.read())
print('N=2 train_parallel_views=$v ms', round(r['ms_per_step'], 3), 'fps', round(r['value'], 1), 'exposed', r['allreduce_ms_exposed_raw'], 'loss', r['final_loss'])"
done
