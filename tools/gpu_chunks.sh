#!/bin/bash
for c in 2 4 8 64; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --chunk-mib $c 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('chunk $c MiB: device ms/step %.3f  e2e ms/step %.3f e2e %.3e'%(d['ms_per_step'],d['e2e']['ms_per_step'],d['e2e']['value']))"
done
