#!/bin/bash
# Round 2, call 1: run every gated test, A/B the three prepared-but-unrun paths, first training/feeder numbers, conv layers vs cuDNN.
O=gpurun_out/r02c01; mkdir -p $O
export DPFT_EXPERIMENTAL=1
timeout 600 python -m pytest tests -m gpu -q --timeout 180 -p no:cacheprovider > $O/pytest_experimental.txt 2>&1
tail -40 $O/pytest_experimental.txt
unset DPFT_EXPERIMENTAL
ab() {  # name, env assignment, extra args
  env $2 timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline $3 2>$O/ab_$1.err | tail -1 > $O/ab_$1.json
  python - "$1" $O/ab_$1.json <<'PY'
import sys, json
try:
    r = json.loads(open(sys.argv[2]).read())
    print(sys.argv[1], 'ms', round(r['ms_per_step'], 4), 'e2e_ms', round(r['e2e']['ms_per_step'], 4), 'seq_ms', round(r['sequential']['ms_per_step'], 4), 'feeder', r.get('e2e_feeder'))
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
}
ab base DPFT_X=0
ab fpn_build2 DPFT_FPN_BUILD=2
ab head16 DPFT_HEAD_LANES=16
ab fpn_fork DPFT_FPN_FORK=1
ab all3 "DPFT_FPN_BUILD=2 DPFT_HEAD_LANES=16 DPFT_FPN_FORK=1"
ab feeder DPFT_X=0 --feeder
DPFT_FPN_BUILD=2 timeout 300 python -m pytest tests/test_features_gpu.py tests/test_model_gpu.py -m gpu -q -x --timeout 120 -p no:cacheprovider 2>&1 | tail -3
DPFT_HEAD_LANES=16 DPFT_FPN_FORK=1 timeout 300 python -m pytest tests/test_model_gpu.py tests/test_infer_stream_gpu.py -m gpu -q -x --timeout 120 -p no:cacheprovider 2>&1 | tail -3
timeout 300 python bench.py --mode train --steps 10 --warmup 3 2>$O/train.err | tail -1 | tee $O/train_dp1.json | cut -c1-300
timeout 300 python bench.py --mode train --steps 10 --warmup 3 --criterion --lsap device 2>$O/train_crit.err | tail -1 | tee $O/train_crit_device.json | cut -c1-300
timeout 300 python tools/conv_bench.py > $O/conv_layers_vs_cudnn.jsonl 2>$O/conv_bench.err; cat $O/conv_layers_vs_cudnn.jsonl | cut -c1-400
