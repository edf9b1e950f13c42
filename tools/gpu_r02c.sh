#!/bin/bash
# fused kernel: full test run + ncu full capture of the kp_fused launches of one step
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $OUT/pytest_gpu_r02c.log
echo "== ncu full (kp_fused)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^kp_fused$' -s 18 -c 6 \
    -o $OUT/prof_r02c -f python bench.py --steps 1 --warmup 3 --no-cpu --no-parity > $OUT/ncu_full_r02c.log 2>&1
tail -3 $OUT/ncu_full_r02c.log
ls -la $OUT | tail -5
