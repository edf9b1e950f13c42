#!/bin/bash
# quick GPU iteration: parity tests + short bench (no CPU baseline leg)
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu_$TAG.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | tee $OUT/bench_$TAG.json
python - <<PY
import json
d=json.load(open("$OUT/bench_$TAG.json"))
print("value %.3e e2e %.3e ms/step %.3f stages %s roofline %.3f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],{k:round(v,3) for k,v in d["stages_ms_per_step"].items()},d["roofline"]["frac"]))
PY
