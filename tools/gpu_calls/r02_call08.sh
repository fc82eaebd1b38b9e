#!/bin/bash
# Round 2, call 8: cap on the persistent conv grid of the side (radar) views: parity, then A/B of the cap on the bench workload.
O=gpurun_out/r02c08; mkdir -p $O
timeout 300 python -m pytest tests/test_conv_gpu.py -m gpu -q --timeout 200 -p no:cacheprovider -k "cta_budget or expand" 2>&1 | tail -4
for n in 0 16 32 48 74; do
DPFT_SIDE_VIEW_CTAS=$n timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-train --no-library-baseline 2>/dev/null | tail -1 > $O/bench_side$n.json
python - $O/bench_side$n.json $n <<'PY'
import sys, json
r = json.load(open(sys.argv[1]))
print('side_view_ctas', sys.argv[2], 'ms', round(r['ms_per_step'], 4), 'e2e', round(r['e2e']['ms_per_step'], 4), 'seq', round(r['sequential']['ms_per_step'], 4), 'sustained', round(r['sustained']['ms_per_step'], 4), 'clk', r['clocks']['sm_mhz'])
PY
done | tee $O/side_view_ctas_ab.txt
