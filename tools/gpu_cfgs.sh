#!/bin/bash
# BASELINE.json configs[2] (1M Wikipedia-shape sentences) and configs[3] (4096 x 4096 chars) at full size, with the
# parity gate on: bench lines for profiles/ (these are parity cases, not the headline workload)
mkdir -p gpurun_out
for w in cfg4 cfg3; do
  timeout 1500 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu 2> gpurun_out/bench_$w.err | tail -1 > gpurun_out/bench_$w.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$w.json"))
    print("$w value %.3e e2e %.3e ms/step %.3f bytes %d parity %s stages %s"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["config"]["bytes_per_gpu"],d["parity"].get("match"),{k[:-3]:round(v,3) for k,v in d["stages_ms_per_step"].items()}))
    print("   counters", d["counters"])
except Exception as e:
    print("$w failed", e, open("gpurun_out/bench_$w.err").read()[-800:])
PY
done
