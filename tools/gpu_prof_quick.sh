#!/bin/bash
# ncu --set full capture (with source) of every kp_ kernel of the third pass of tools/quick_bench.py
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'kp_' -s 46 -c 20 \
    -o gpurun_out/prof_$TAG -f python tools/quick_bench.py cfg2 65536 > gpurun_out/ncu_full_$TAG.log 2>&1
tail -4 gpurun_out/ncu_full_$TAG.log | cut -c1-200
ls -la gpurun_out/prof_$TAG.ncu-rep
