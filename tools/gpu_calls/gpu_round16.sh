#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/conv_debug.py pair_pw_small pair_pw_tail pair_c3 pair_big_pw pair_big_c3 big_pw big_c3 > gpurun_out/conv_debug.jsonl 2> gpurun_out/conv_debug.err
cut -c1-400 gpurun_out/conv_debug.jsonl; tail -3 gpurun_out/conv_debug.err
