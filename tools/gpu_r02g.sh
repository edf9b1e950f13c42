#!/bin/bash
OUT=gpurun_out
TAG=${1:-r02g}
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu_$TAG.log
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu 2> $OUT/bench_$TAG.err | tail -1 | tee $OUT/bench_$TAG.json | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print({k:j[k] for k in ('value','ms_per_step','stages_ms_per_step','gpu_launches')}); print(j['parity']['match'], 'e2e', j['e2e']['value'], j['e2e']['sync_call']['value'])"
tail -3 $OUT/bench_$TAG.err
