#!/bin/bash
mkdir -p gpurun_out
for w in cfg4 cfg3; do
  timeout 900 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/bench_$w.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$w.json"))
    print("$w value %.3e e2e %.3e ms/step %.3f bytes %d stages %s"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["config"]["bytes_per_gpu"],{k[:-3]:round(v,3) for k,v in d["stages_ms_per_step"].items()}))
    print("   counters", d["counters"])
except Exception as e:
    print("$w failed", e, open("gpurun_out/bench_$w.json").read()[-600:])
PY
done
