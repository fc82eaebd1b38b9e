#!/bin/bash
# Round 2, call 10: L2 prefetch pass of the decoder gather (parity + A/B), pipeline depth with the side-view cap.
O=gpurun_out/r02c10; mkdir -p $O
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_golden_taps_gpu.py tests/test_infer_stream_gpu.py tests/test_reference_on_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider -x 2>&1 | tail -4
run() {  # label, env, args
  env $2 timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-train --no-library-baseline $3 2>/dev/null | tail -1 > $O/b.json
  python - $O/b.json "$1" <<'PY'
import sys, json
r = json.load(open(sys.argv[1]))
print(sys.argv[2], 'ms', round(r['ms_per_step'], 4), 'e2e', round(r['e2e']['ms_per_step'], 4), 'seq', round(r['sequential']['ms_per_step'], 4), 'sustained', round(r['sustained']['ms_per_step'], 4), 'dec us', round(r['roofline_decoder']['us_per_launch'], 1), 'clk', r['clocks']['sm_mhz'])
PY
}
{
run "prefetch=0 depth=3" DPFT_DECODER_PREFETCH=0 ""
run "prefetch=1 depth=3" DPFT_DECODER_PREFETCH=1 ""
run "prefetch=0 depth=3 (again)" DPFT_DECODER_PREFETCH=0 ""
run "prefetch=1 depth=3 (again)" DPFT_DECODER_PREFETCH=1 ""
run "prefetch=1 depth=2" DPFT_DECODER_PREFETCH=1 "--depth 2"
run "prefetch=1 depth=4" DPFT_DECODER_PREFETCH=1 "--depth 4"
} | tee $O/decoder_prefetch_depth_ab.txt
