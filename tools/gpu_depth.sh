#!/bin/bash
OUT=gpurun_out
for d in 1 2 3 4; do
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-parity --queue-depth $d 2>/dev/null | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('depth', $d, 'device', round(j['value']/1e9,2), 'e2e', round(j['e2e']['value']/1e9,2), 'ms', round(j['e2e']['ms_per_step'],3), 'sync', round(j['e2e']['sync_call']['value']/1e9,2))"
done
