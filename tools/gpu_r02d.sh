#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
echo "== fused tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "both_paths or fused_path or cfg1 or ragged or full_cfg2 or long_unknown or first_character" 2>&1 | tail -30 | tee $OUT/pytest_fused_r02d.log
echo "== bench auto"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu 2> $OUT/bench_r02d.err | tail -1 | tee $OUT/bench_r02d.json | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print({k:j[k] for k in ('value','ms_per_step','parity','stages_ms_per_step','fused_sentences_per_step','gpu_launches')}); print(j['e2e']['value'], j['e2e']['sync_call']['value'])"
tail -3 $OUT/bench_r02d.err
