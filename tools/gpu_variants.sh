#!/bin/bash
# bench every library variant under kanpyo_b200/_variants (perf only; parity is checked separately)
mkdir -p gpurun_out
for so in kanpyo_b200/_variants/libkanpyo_b200.*.so; do
  n=$(basename $so | sed 's/libkanpyo_b200\.\(.*\)\.so/\1/')
  KANPYO_B200_LIB=$PWD/$so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/var_$n.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/var_$n.json"))
    print("%-10s ms/step %.3f stages %s"%("$n",d["ms_per_step"],{k[:-3]:round(v,3) for k,v in d["stages_ms_per_step"].items()}))
except Exception as e:
    print("$n failed", e, open("gpurun_out/var_$n.json").read()[:300])
PY
done
