#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu_r02e.log
echo "== latency sweep"; timeout 900 python tools/latency_sweep.py > $OUT/latency_sweep_r02e.json 2> $OUT/latency_sweep_r02e.err; cat $OUT/latency_sweep_r02e.err | tail -14
