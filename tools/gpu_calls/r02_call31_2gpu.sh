#!/bin/bash
# Round 2, call 31 (2 GPUs): the default bench line at N=2 twice, tight timeouts.
O=gpurun_out/r02c31; mkdir -p $O
for rep in 1; do
T0=$(date +%s)
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29610 + rep)) bench.py --gpus 2 --steps 20 --warmup 5 --train-timeout 120 > $O/bench_n2_$rep.out 2> $O/bench_n2_$rep.err
echo "bench N=2 rep $rep rc=$? wall=$(( $(date +%s) - T0 )) s"
tail -1 $O/bench_n2_$rep.out > $O/bench_n2_$rep.json
python - $rep <<'PY'
import json, sys
r = json.load(open(f'gpurun_out/r02c31/bench_n2_{sys.argv[1]}.json'))
t = r['train']
print({k: r[k] for k in ('value', 'ms_per_step')}, 'e2e', round(r['e2e']['value'], 1), 'train', {k: t.get(k) for k in ('value', 'ms_per_step', 'allreduce_ms_exposed_raw', 'final_loss', 'error')})
PY
done
