#!/bin/bash
# eight GPUs: the torchrun bench (weak line + strong sub-line, gathers inside the timed steps) and the single-process shards test
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name --format=csv | tail -8 | sort | uniq -c
echo "== bench N=8"; timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 2> $OUT/bench_r02h_n8.err | tail -1 | tee $OUT/bench_r02h_n8.json | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print({k:j.get(k) for k in ('value','ms_per_step','parity','token_gather_ms','dict_broadcast_ms','strong_scaling')}); print('e2e',j['e2e']['value'], j['e2e']['sync_call']['value'])"
tail -3 $OUT/bench_r02h_n8.err
echo "== shards test (8 devices, one process)"; timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "shards" 2>&1 | tail -5 | tee $OUT/pytest_shards_r02h.log
