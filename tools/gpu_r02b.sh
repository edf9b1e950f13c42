#!/bin/bash
# fused kernel bring-up: the path tests first (fail fast), then everything, then bench lines for both paths
OUT=gpurun_out
mkdir -p $OUT
echo "== fused tests"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "both_paths or fused_path or cfg1 or ragged" 2>&1 | tail -30 | tee $OUT/pytest_fused_r02b.log
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $OUT/pytest_gpu_r02b.log
echo "== bench auto"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu 2> $OUT/bench_r02b.err | tail -1 | tee $OUT/bench_r02b.json
tail -5 $OUT/bench_r02b.err
echo "== bench pipeline"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --path pipeline 2> $OUT/bench_r02b_pipe.err | tail -1 | tee $OUT/bench_r02b_pipe.json
