#!/bin/bash
nproc
for b in off on; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2951$( [ $b = on ] && echo 3 || echo 4 ) bench.py --gpus 8 --steps 10 --warmup 3 --no-parity --blocking-sync $b 2>/dev/null | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('blocking', '$b', 'device', round(j['value']/1e9,2), 'e2e', round(j['e2e']['value']/1e9,2), 'ms', round(j['e2e']['ms_per_step'],3), 'sync', round(j['e2e']['sync_call']['value']/1e9,2), j['e2e']['queue_workers_wait'], j['config'].get('host_affinity'))"
done
