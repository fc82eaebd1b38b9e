#!/bin/bash
# Round 2, call 7: whole GPU suite with the new defaults and bars; launch list + conv traffic of one eval step under ncu;
# ncu --set full of the decoder layer kernel and of the FPN raw-level kernel; default bench line with its wall time.
O=gpurun_out/r02c07; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > $O/pytest_gpu.txt 2>&1
tail -12 $O/pytest_gpu.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_eval_step.csv python tools/one_forward.py 2 > $O/ncu_launches.log 2>&1
tail -1 $O/ncu_launches.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --profile-from-start off -k 'regex:conv_gemm|conv3x3_halo|conv_expand' --csv --log-file $O/conv_traffic.csv python tools/one_forward.py 2 > $O/ncu_conv_traffic.log 2>&1
tail -1 $O/ncu_conv_traffic.log
python tools/ncu_summaries.py $O/launches_eval_step.csv $O/conv_traffic.csv $O/r02
timeout 300 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:decoder_layer -c 1 -f -o $O/ncu_decoder_layer python tools/one_forward.py 2 > $O/ncu_decoder.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:fpn_output_tc2 -c 1 -f -o $O/ncu_fpn_tc2 python tools/one_forward.py 2 > $O/ncu_fpn.log 2>&1
T0=$(date +%s)
timeout 900 python bench.py --steps 20 --warmup 5 2>$O/bench.err | tail -1 > $O/bench_default.json
echo "bench wall $(( $(date +%s) - T0 )) s"; tail -2 $O/bench.err
python - <<'PY'
import json
r = json.load(open('gpurun_out/r02c07/bench_default.json'))
for k in ('value', 'ms_per_step', 'e2e', 'sequential', 'sustained', 'clocks', 'cpu_baseline'):
    print(k, json.dumps(r.get(k))[:600])
print('parity', json.dumps({k: v for k, v in r['parity'].items() if k != 'outputs'})[:900])
print('lib', json.dumps(r['gpu_library_baseline'])[:400])
print('train', r['train']['ms_per_step'], 'roof', r['roofline']['frac'], r['roofline']['in_step'], 'dec', r['roofline_decoder']['frac'], r['roofline_decoder']['us_per_launch'])
PY
ls -la $O
