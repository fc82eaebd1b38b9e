#!/bin/bash
# Round 2, call 15: two equal-shaped side views in one launch per Bottleneck convolution: parity, then A/B on the bench workload
# (pairing on/off x persistent-grid cap of the side views).
O=gpurun_out/r02c15; mkdir -p $O
timeout 400 python -m pytest tests/test_conv_gpu.py tests/test_infer_stream_gpu.py tests/test_model_gpu.py -m gpu -q --timeout 200 -p no:cacheprovider -x 2>&1 | tail -4
run() {
  env $2 timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-train --no-library-baseline 2>/dev/null | tail -1 > $O/b.json
  python - $O/b.json "$1" <<'PY'
import sys, json
r = json.load(open(sys.argv[1]))
print(sys.argv[2], 'ms', round(r['ms_per_step'], 4), 'e2e', round(r['e2e']['ms_per_step'], 4), 'seq', round(r['sequential']['ms_per_step'], 4), 'sustained', round(r['sustained']['ms_per_step'], 4), 'roof', round(r['roofline']['frac'], 4), 'kernel ms', round(r['roofline']['ms_in_kernel_per_step'], 3), 'launches', r['roofline']['launches'], 'gpu_launches', r['gpu_launches'], 'clk', r['clocks']['sm_mhz'])
PY
}
{
run "pair=0 cap=48" "DPFT_PAIR_SIDE_VIEWS=0 DPFT_SIDE_VIEW_CTAS=48"
run "pair=1 cap=48" "DPFT_PAIR_SIDE_VIEWS=1 DPFT_SIDE_VIEW_CTAS=48"
run "pair=1 cap=64" "DPFT_PAIR_SIDE_VIEWS=1 DPFT_SIDE_VIEW_CTAS=64"
run "pair=1 cap=96" "DPFT_PAIR_SIDE_VIEWS=1 DPFT_SIDE_VIEW_CTAS=96"
run "pair=0 cap=48 (again)" "DPFT_PAIR_SIDE_VIEWS=0 DPFT_SIDE_VIEW_CTAS=48"
run "pair=1 cap=64 (again)" "DPFT_PAIR_SIDE_VIEWS=1 DPFT_SIDE_VIEW_CTAS=64"
} | tee $O/pair_side_views_ab.txt
