#!/bin/bash
# Round 2, call 11 (4 GPUs): the default bench line at N=4 exactly as the driver launches it; tight timeouts.
O=gpurun_out/r02c11; mkdir -p $O
nvidia-smi -L | wc -l
T0=$(date +%s)
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 20 --warmup 5 --train-timeout 120 > $O/bench_n4.out 2> $O/bench_n4.err
echo "bench N=4 rc=$? wall=$(( $(date +%s) - T0 )) s"
tail -1 $O/bench_n4.out > $O/bench_n4.json; grep -v "OMP_NUM_THREADS\|\*\*\*\*" $O/bench_n4.err | tail -5
python - <<'PY'
import json
try:
    r = json.load(open('gpurun_out/r02c11/bench_n4.json'))
    print({k: r[k] for k in ('value', 'ms_per_step', 'n_gpus')}, 'e2e', r['e2e']['value'], r['e2e']['ms_per_step'])
    t = r['train']
    print('train', {k: t.get(k) for k in ('value', 'ms_per_step', 'allreduce_ms_exposed', 'ms_per_step_without_allreduce', 'n_gpus', 'error')})
except Exception as e:
    print('FAILED to parse', e)
PY
