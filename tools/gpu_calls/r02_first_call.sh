#!/bin/bash
# First gpurun call of round 2: validate what was written after round 1's GPU budget ran out (all of it is off by default).
#   gpurun --timeout 900 -- 'bash tools/gpu_calls/r02_first_call.sh > gpurun_out/r02_first_call.log 2>&1'
# 1. experimental tests (column-owning FPN raw-level builder = dpft_fpn_output_forward impl 3; golden taps / gradient digests
#    of the reference on the GPU paths)
export DPFT_EXPERIMENTAL=1
timeout 120 python -m pytest tests/test_features_gpu.py -m gpu -q -x --timeout 60 -k "column_builder" 2>&1 | tail -5
timeout 300 python -m pytest tests/test_golden_taps_gpu.py -m gpu -q --timeout 120 2>&1 | tail -8
timeout 120 python -m pytest tests/test_decoder_head16_gpu.py -m gpu -q -x --timeout 60 2>&1 | tail -5
timeout 120 python -m pytest tests/test_infer_stream_gpu.py -m gpu -q -x --timeout 60 -k feeder 2>&1 | tail -5
timeout 120 python -m pytest tests/test_criterion_metrics_gpu.py -m gpu -q --timeout 60 2>&1 | tail -5
unset DPFT_EXPERIMENTAL
# 2. the whole model with the experimental builder chosen by the automatic path, then the A/B on the bench workload
DPFT_FPN_BUILD=2 timeout 300 python -m pytest tests/test_features_gpu.py tests/test_model_gpu.py -m gpu -q -x --timeout 120 2>&1 | tail -3
for v in 1 2; do
DPFT_FPN_BUILD=$v timeout 200 python bench.py --steps 20 --warmup 4 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys,json
r=json.loads(sys.stdin.read()); print('fpn_build=$v', 'ms', r['ms_per_step'], 'fps', r['value'], 'e2e', r['e2e']['value'], 'seq', r['sequential']['ms_per_step'])"
done
# 2b. sixteen-lanes-per-query head kernel (bit-identical by construction): whole-model tests, then the A/B
DPFT_HEAD_LANES=16 timeout 300 python -m pytest tests/test_model_gpu.py -m gpu -q -x --timeout 120 2>&1 | tail -3
for v in 1 16; do
DPFT_HEAD_LANES=$v timeout 200 python bench.py --steps 20 --warmup 4 --no-cpu-baseline --depth 1 2>/dev/null | tail -1 | python -c "
import sys,json
r=json.loads(sys.stdin.read()); print('head_lanes=$v', 'sequential ms', r['ms_per_step'])"
done
# 2c. forked FPN output launches (same kernels, different order / stream): model tests, then the A/B
DPFT_FPN_FORK=1 timeout 300 python -m pytest tests/test_features_gpu.py tests/test_model_gpu.py tests/test_infer_stream_gpu.py -m gpu -q -x --timeout 120 2>&1 | tail -3
for v in 0 1; do
DPFT_FPN_FORK=$v timeout 200 python bench.py --steps 20 --warmup 4 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys,json
r=json.loads(sys.stdin.read()); print('fpn_fork=$v', 'ms', r['ms_per_step'], 'seq', r['sequential']['ms_per_step'])"
done
# 3. per-stage times with the experimental builder (camera_mono.pyramid_total - camera_mono.backbone = the FPN part)
DPFT_FPN_BUILD=2 timeout 200 python tools/stage_times.py 2>/dev/null | tail -1
# 4. e2e through the batch feeder (uint8 camera frames uploaded, dataset arithmetic on the GPU): compare e2e_feeder with e2e
timeout 200 python bench.py --steps 20 --warmup 4 --no-cpu-baseline --feeder 2>/dev/null | tail -1 | python -c "
import sys,json
r=json.loads(sys.stdin.read()); print('e2e', r['e2e'], 'e2e_feeder', r['e2e_feeder'])"
# 5. training step with the reference criterion (eager: the assignment is a host solve) next to the fixed scalar loss
timeout 300 python bench.py --mode train --steps 10 --warmup 3 --criterion 2>/dev/null | tail -1
timeout 300 python bench.py --mode train --steps 10 --warmup 3 --criterion --lsap device 2>/dev/null | tail -1   # one graph per step
timeout 300 python bench.py --mode train --steps 10 --warmup 3 --no-graph 2>/dev/null | tail -1
