#!/bin/bash
# launch list + full capture of every kp_ kernel of one step
TAG=${1:-all}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_launches_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'kp_' -s 40 -c 20 \
    -o gpurun_out/prof_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log | cut -c1-200
