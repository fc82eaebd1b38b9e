#!/bin/bash
# One GPU-box session: parity tests, smoke, bench, MSDA sweep, ncu launch list + full capture of the top kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 300 python tools/msda_sweep.py > gpurun_out/msda_sweep.jsonl 2> gpurun_out/msda_sweep.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:msda_fwd -s 3 -c 2 -o gpurun_out/msda_fwd \
    python tools/msda_sweep.py --only cfg5_bf16_D32 --reps 2 > gpurun_out/ncu_msda.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:msda_bwd -s 3 -c 2 -o gpurun_out/msda_bwd \
    python tools/msda_sweep.py --only cfg5_bf16_D32 --reps 2 >> gpurun_out/ncu_msda.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -2; cat gpurun_out/msda_sweep.jsonl; cat gpurun_out/bench.json
