#!/bin/bash
# Round 2, call 12: uint8 camera frames read in place by the stem / FPN kernels (parity, e2e), H2D probe, BN column-sum change
# (training tests + step time), whole suite.
O=gpurun_out/r02c12; mkdir -p $O
timeout 120 python tools/h2d_probe.py 2>&1 | tail -1 | tee $O/h2d_probe.json
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > $O/pytest_gpu.txt 2>&1
tail -8 $O/pytest_gpu.txt
T0=$(date +%s)
timeout 900 python bench.py --steps 20 --warmup 5 2>$O/bench.err | tail -1 > $O/bench_default.json
echo "bench wall $(( $(date +%s) - T0 )) s"; tail -2 $O/bench.err
python - <<'PY'
import json
r = json.load(open('gpurun_out/r02c12/bench_default.json'))
for k in ('value', 'ms_per_step', 'e2e', 'e2e_fp32_inputs', 'sequential', 'sustained', 'clocks'):
    print(k, json.dumps(r.get(k))[:500])
print('parity', json.dumps({k: v for k, v in r['parity'].items() if k != 'outputs'})[:700])
print('train', r['train']['ms_per_step'], 'roof', r['roofline']['frac'], r['roofline']['in_step']['frac_lower_bound'])
PY
timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-train --no-library-baseline 2>/dev/null | tail -1 | python -c "
import sys, json
r = json.loads(sys.stdin.read())
print('60 steps: value ms', r['ms_per_step'], 'e2e ms', r['e2e']['ms_per_step'], 'e2e fp32 ms', r['e2e_fp32_inputs']['ms_per_step'], 'seq', r['sequential']['ms_per_step'])"
