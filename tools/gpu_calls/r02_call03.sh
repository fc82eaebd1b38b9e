#!/bin/bash
# Round 2, call 3: weight-stationary expand kernel (parity + per-layer A/B + bench A/B), stream-tuned CTA pairs on the expand
# layers, parity diagnostics at full size with the TF32 yardstick, the re-based gradient digest test.
O=gpurun_out/r02c03; mkdir -p $O
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_golden_taps_gpu.py tests/test_full_size_gpu.py tests/test_reference_on_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider -x > $O/pytest.txt 2>&1
tail -15 $O/pytest.txt
for ws in 0 1; do
DPFT_CONV_WS=$ws timeout 200 python tools/conv_bench.py s1_conv3 s2_conv3 s3_conv3 s4_conv3 2>&1 | cut -c1-330 | sed "s/^/ws=$ws /"
done | tee $O/conv_expand_ws_ab.txt
timeout 200 python tools/conv_bench.py --sweep s1_conv3 s2_conv3 s3_conv3 s4_conv3 2>&1 | tee $O/conv_expand_sweep.txt
for ws in 0 1; do
DPFT_CONV_WS=$ws timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-train --no-library-baseline 2>/dev/null | tail -1 > $O/bench_ws$ws.json
python - $O/bench_ws$ws.json $ws <<'PY'
import sys, json
r = json.load(open(sys.argv[1]))
print('ws', sys.argv[2], 'ms', round(r['ms_per_step'], 4), 'seq', round(r['sequential']['ms_per_step'], 4), 'roof', round(r['roofline']['frac'], 4), 'kernel ms', round(r['roofline']['ms_in_kernel_per_step'], 3))
PY
done
timeout 600 python tools/parity_at_size.py config2 config3 2>/dev/null | tee $O/parity_at_size.jsonl | cut -c1-400
