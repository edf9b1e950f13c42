#!/bin/bash
# round 2, first GPU call: parity tests (new full-size cases, compact records, queue, shards) + one bench line
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total --format=csv | tail -1; nproc
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $OUT/pytest_gpu_r02a.log
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 2> $OUT/bench_r02a.err | tail -1 | tee $OUT/bench_r02a.json
tail -5 $OUT/bench_r02a.err
