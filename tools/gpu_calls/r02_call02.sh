#!/bin/bash
# Round 2, call 2: whole GPU suite without any gate (new: full-size parity, reference on GPU, reference-init cases), the new
# default bench line (parity / train / gpu_library_baseline / sustained), the reference arm, ncu of the two stage-3 1x1 layers.
O=gpurun_out/r02c02; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > $O/pytest_gpu.txt 2>&1
tail -25 $O/pytest_gpu.txt
timeout 600 python bench.py 2>$O/bench.err | tail -1 > $O/bench_default.json; tail -3 $O/bench.err
python - <<'PY'
import json
r = json.load(open('gpurun_out/r02c02/bench_default.json'))
for k in ('value', 'ms_per_step', 'e2e', 'sequential', 'sustained', 'clocks', 'cpu_baseline', 'parity', 'gpu_library_baseline', 'train'):
    print(k, json.dumps(r.get(k))[:1500])
print('roofline', {k: r['roofline'][k] for k in ('frac', 'achieved', 'ms_in_kernel_per_step', 'frac_of_attainable', 'in_step')})
PY
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-700
for L in s3_conv3 s3_conv1; do
timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_gemm -c 2 -f -o $O/ncu_$L python tools/conv_bench.py $L > $O/ncu_$L.log 2>&1
done
ls -la $O
