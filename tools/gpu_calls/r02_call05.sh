#!/bin/bash
# Round 2, call 5: more weight-stationary splits; operand stages instead of the unused residual ring on the deep-K layers;
# bench A/B of the best; whole-bench wall time.
O=gpurun_out/r02c05; mkdir -p $O
for v in 3 6 7; do
DPFT_WS_VARIANT=$v timeout 120 python tools/conv_bench.py --no-lib s3_conv3 2>&1 | cut -c1-120 | sed "s/^/wsv=$v /"
done | tee $O/ws_variants2.txt
for v in 3 8 9; do
DPFT_WS_VARIANT=$v timeout 120 python tools/conv_bench.py --no-lib s2_conv3 2>&1 | cut -c1-120 | sed "s/^/wsv=$v /"
done | tee -a $O/ws_variants2.txt
for v in 0 1; do
DPFT_CONV_DEEP_VARIANT=$v timeout 120 python tools/conv_bench.py --no-lib s2_conv2 s3_conv1 s3_conv2 s4_conv1 s4_conv2 s3_conv2_s2 2>&1 | cut -c1-120 | sed "s/^/deep=$v /"
done | tee $O/deep_variants.txt
for cfgv in "DPFT_WS_VARIANT=0 DPFT_CONV_DEEP_VARIANT=0" "DPFT_WS_VARIANT=3 DPFT_CONV_DEEP_VARIANT=0" "DPFT_WS_VARIANT=3 DPFT_CONV_DEEP_VARIANT=1"; do
env $cfgv timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-train --no-library-baseline 2>/dev/null | tail -1 > $O/bench_ab.json
python - $O/bench_ab.json "$cfgv" <<'PY'
import sys, json
r = json.load(open(sys.argv[1]))
print(sys.argv[2], 'ms', round(r['ms_per_step'], 4), 'seq', round(r['sequential']['ms_per_step'], 4), 'roof', round(r['roofline']['frac'], 4), 'kernel ms', round(r['roofline']['ms_in_kernel_per_step'], 3), 'dec', r['roofline_decoder']['us_per_launch'] if r.get('roofline_decoder') else None)
PY
done | tee $O/bench_ab.txt
