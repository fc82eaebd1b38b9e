#!/bin/bash
# Round 2, call 9 (8 GPUs): the default bench line at N=8 exactly as the driver launches it (inference replicas + the config-4
# training step with its NCCL all-reduces inside the captured graph), and at N=4; wall time of each.
O=gpurun_out/r02c09; mkdir -p $O
nvidia-smi -L | wc -l
for N in 8; do
T0=$(date +%s)
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520 + N)) bench.py --gpus $N --steps 20 --warmup 5 --train-timeout 100 > $O/bench_n$N.out 2> $O/bench_n$N.err
echo "bench N=$N rc=$? wall=$(( $(date +%s) - T0 )) s"
tail -1 $O/bench_n$N.out > $O/bench_n$N.json; grep -v "OMP_NUM_THREADS\|\*\*\*\*" $O/bench_n$N.err | tail -5
python - $N <<'PY'
import json, sys
N = sys.argv[1]
try:
    r = json.load(open(f'gpurun_out/r02c09/bench_n{N}.json'))
    print({k: r[k] for k in ('value', 'ms_per_step', 'n_gpus')}, 'e2e', r['e2e']['value'], r['e2e']['ms_per_step'])
    t = r['train']
    print('train', {k: t[k] for k in ('value', 'ms_per_step', 'allreduce_ms_exposed', 'ms_per_step_without_allreduce', 'n_gpus')})
except Exception as e:
    print('FAILED to parse', e)
PY
done
