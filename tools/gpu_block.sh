#!/bin/bash
for b in off on; do
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-parity --blocking-sync $b 2>/dev/null | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('blocking', '$b', 'device', round(j['value']/1e9,2), 'e2e', round(j['e2e']['value']/1e9,2), 'ms', round(j['e2e']['ms_per_step'],3), j['e2e']['queue_workers_wait'])"
done
