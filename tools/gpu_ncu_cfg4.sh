#!/bin/bash
# ncu --set full of the long-line workload (BASELINE configs[3]: 4096 x 4096 chars): sweep, back-trace, bucketize, lattice kernels
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'kp_viterbi|kp_lattice_walk|kp_lattice_count|kp_bucketize|kp_backtrace_find' -s 15 -c 5 \
    -o $OUT/prof_cfg4 -f python bench.py --workload cfg4 --steps 1 --warmup 3 --no-cpu --no-parity > $OUT/ncu_full_cfg4.log 2>&1
tail -2 $OUT/ncu_full_cfg4.log | cut -c1-200
