#!/bin/bash
# compute-sanitizer over both device paths (fused kernel: shared-memory races matter most)
TAG=${1:-x}
OUT=gpurun_out
mkdir -p $OUT
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer $tool"
  timeout 1200 compute-sanitizer --tool $tool --target-processes all --launch-timeout 300 python tools/small_case.py 40 > $OUT/${tool}_$TAG.log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|PARITY|hazard" $OUT/${tool}_$TAG.log | sort | uniq -c | head -8
done
