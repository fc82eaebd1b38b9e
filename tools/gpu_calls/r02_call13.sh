#!/bin/bash
# Round 2, call 13: spare pipeline slots for host batches (upload overlaps running forwards): parity + e2e A/B.
O=gpurun_out/r02c13; mkdir -p $O
timeout 300 python -m pytest tests/test_infer_stream_gpu.py tests/test_model_gpu.py -m gpu -q --timeout 200 -p no:cacheprovider 2>&1 | tail -3
for n in 0 1 2 3; do
DPFT_UPLOAD_SLOTS=$n timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-train --no-library-baseline 2>/dev/null | tail -1 | python -c "
import sys, json
r = json.loads(sys.stdin.read())
print('upload_slots=$n value ms', round(r['ms_per_step'], 4), 'e2e(u8) ms', round(r['e2e']['ms_per_step'], 4), 'e2e fp32 ms', round(r['e2e_fp32_inputs']['ms_per_step'], 4), 'seq', round(r['sequential']['ms_per_step'], 4), 'clk', r['clocks']['sm_mhz'])"
done | tee $O/upload_slots_ab.txt
