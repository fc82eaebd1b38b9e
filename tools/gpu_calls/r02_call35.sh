#!/bin/bash
# Round 2, call 35: ncu --set full of the weight-stationary expand kernel (stage-3 conv3) and of the FPN raw-level kernel.
O=gpurun_out/r02c35; mkdir -p $O
timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_expand_ws -c 2 -f -o $O/ncu_ws_s3_conv3 python tools/conv_bench.py --no-lib s3_conv3 > $O/ncu_ws.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:fpn_output_tc2 -c 1 -f -o $O/ncu_fpn_tc2 python tools/one_forward.py 2 > $O/ncu_fpn.log 2>&1
ls -la $O
