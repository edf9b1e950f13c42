#!/bin/bash
# ncu --set full capture of kernels matching $1 (regex), $2 = tag, $3 = skip count, $4 = count
RE=${1:-kp_viterbi}; TAG=${2:-x}; SKIP=${3:-3}; CNT=${4:-1}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $SKIP -c $CNT \
    -o gpurun_out/prof_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_$TAG.log 2>&1
tail -3 gpurun_out/ncu_$TAG.log | cut -c1-300
