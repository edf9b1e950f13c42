#!/bin/bash
# two GPUs: single-process shards over NCCL (C ABI) + the torchrun bench with the gather inside the timed step
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name --format=csv | tail -2
echo "== shards test"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "shards or sharded" 2>&1 | tail -15 | tee $OUT/pytest_shards_r02f.log
echo "== bench N=2"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2> $OUT/bench_r02f_n2.err | tail -1 | tee $OUT/bench_r02f_n2.json | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print({k:j.get(k) for k in ('value','ms_per_step','parity','token_gather_ms','dict_broadcast_ms','strong_scaling')}); print('e2e',j['e2e']['value'], j['e2e']['sync_call']['value'])"
tail -5 $OUT/bench_r02f_n2.err
