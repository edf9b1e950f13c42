#!/bin/bash
# One gpurun call: parity tests of the in-tree library, then tools/sweep.py over every variant.
# usage: gpurun --timeout 1200 -- 'bash tools/gpu_sweep.sh [steps]'
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/sweep.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tail -1
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $OUT/pytest_gpu_sweep.log
echo "== sweep"; timeout 1500 python tools/sweep.py --steps ${1:-10}
echo "== memcheck (64 sentences + edge cases)"
timeout 600 compute-sanitizer --tool memcheck --target-processes all --launch-timeout 120 python tools/small_case.py 64 > $OUT/sanitize_sweep.log 2>&1
grep -E "ERROR SUMMARY|PARITY|Error" $OUT/sanitize_sweep.log | head -5
