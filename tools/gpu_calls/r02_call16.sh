#!/bin/bash
# Round 2, call 16: FPN raw-level builder with the lateral weights as kernel parameters: parity + A/B (stage times, bench).
O=gpurun_out/r02c16; mkdir -p $O
timeout 400 python -m pytest tests/test_features_gpu.py tests/test_model_gpu.py tests/test_infer_stream_gpu.py -m gpu -q --timeout 200 -p no:cacheprovider -x 2>&1 | tail -4
for v in 0 1; do
DPFT_FPN_LAT_PARAMS=$v timeout 200 python tools/stage_times.py 2>/dev/null | tail -1 | python -c "
import sys, json
r = json.loads(sys.stdin.read())
print('lat_params=$v', {k: round(v, 1) for k, v in r.items() if 'pyramid_total' in k or 'backbone' in k or k.startswith('forward') or k == 'decoder'})"
done | tee $O/fpn_lat_params_ab.txt
for v in 0 1 0 1; do
DPFT_FPN_LAT_PARAMS=$v timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-train --no-library-baseline 2>/dev/null | tail -1 | python -c "
import sys, json
r = json.loads(sys.stdin.read())
print('lat_params=$v ms', round(r['ms_per_step'], 4), 'e2e', round(r['e2e']['ms_per_step'], 4), 'seq', round(r['sequential']['ms_per_step'], 4), 'sustained', round(r['sustained']['ms_per_step'], 4), 'clk', r['clocks']['sm_mhz'])"
done | tee -a $O/fpn_lat_params_ab.txt
