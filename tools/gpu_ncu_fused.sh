#!/bin/bash
# ncu --set full of the kp_fused launches of one warmed-up step -> gpurun_out/prof_$1.ncu-rep
TAG=${1:-x}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^kp_fused$' -s 18 -c 6 \
    -o $OUT/prof_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu --no-parity > $OUT/ncu_full_$TAG.log 2>&1
tail -2 $OUT/ncu_full_$TAG.log | cut -c1-300
