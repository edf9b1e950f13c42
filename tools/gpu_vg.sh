#!/bin/bash
for v in "" vg16; do
  if [ -n "$v" ]; then export KANPYO_B200_LIB=$PWD/kanpyo_b200/_variants/libkanpyo_b200.$v.so; else unset KANPYO_B200_LIB; fi
  timeout 600 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu 2>/dev/null | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('variant', '$v', j['value'], j['ms_per_step'], j['parity'].get('match'), j['stages_ms_per_step'])"
done
