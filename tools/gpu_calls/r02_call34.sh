#!/bin/bash
# Round 2, call 34: final validation — whole GPU suite, smoke(), the default bench line, ncu launch list + conv traffic.
O=gpurun_out/r02c34; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > $O/pytest_gpu.txt 2>&1
tail -4 $O/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
T0=$(date +%s)
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 2>$O/bench.err | tail -1 > $O/bench_default.json
echo "bench wall $(( $(date +%s) - T0 )) s"
python - <<'PY'
import json
r = json.load(open('gpurun_out/r02c34/bench_default.json'))
for k in ('value', 'ms_per_step', 'e2e', 'e2e_fp32_inputs', 'sequential', 'sustained', 'clocks', 'gpu_launches'):
    print(k, json.dumps(r.get(k))[:300])
print('parity', json.dumps({k: v for k, v in r['parity'].items() if k not in ('outputs', 'against', 'yardstick_reference_on_gpu_tf32_default')})[:400])
print('lib', r['gpu_library_baseline']['ms_per_step'], r['gpu_library_baseline']['speedup_of_this_repo'], 'cpu', r['cpu_baseline']['value'])
print('train', r['train']['ms_per_step'], r['train']['value'], 'roof', r['roofline']['frac'], r['roofline']['launches'], r['roofline']['tensor_bound_launches'], r['roofline']['frac_of_attainable'], r['roofline']['in_step']['frac_lower_bound'])
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_eval_step.csv python tools/one_forward.py 2 > $O/ncu_launches.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --profile-from-start off -k 'regex:conv_gemm|conv3x3_halo|conv_expand' --csv --log-file $O/conv_traffic.csv python tools/one_forward.py 2 > $O/ncu_conv_traffic.log 2>&1
python tools/ncu_summaries.py $O/launches_eval_step.csv $O/conv_traffic.csv $O/r02final2
