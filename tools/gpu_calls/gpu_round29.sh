#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:decoder_layer -c 1 -f -o gpurun_out/decoder_layer python tools/one_forward.py 2 > gpurun_out/ncu_dec.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:fpn_output_tc_kernel -c 1 -f -o gpurun_out/fpn_out_cam python tools/one_forward.py 2 > gpurun_out/ncu_fpn.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:stem_tc_kernel -c 1 -f -o gpurun_out/stem_cam python tools/one_forward.py 2 > gpurun_out/ncu_stem.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -n 2 gpurun_out/ncu_dec.log
