#!/bin/bash
# per-kernel durations (ncu, serialised, cold cache: compare shares) of the third pass of tools/quick_bench.py
# usage: bash tools/gpu_launches.sh <tag> [lib.so]
TAG=${1:-l}
[ -n "$2" ] && export KANPYO_B200_LIB=$PWD/$2
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'kp_' -s 50 -c 24 --csv --log-file gpurun_out/launches_$TAG.csv \
    python tools/quick_bench.py cfg2 65536 > gpurun_out/ncu_launches_$TAG.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches_$TAG.csv")) if len(r)>5]
h=rows[0]; kn=h.index("Kernel Name"); mv=h.index("Metric Value")
for r in rows[1:]:
    print("%-40s %8.1f us"%(r[kn].split("(")[0][:40], float(r[mv].replace(",",""))/1000 if float(r[mv].replace(",",""))>5000 else float(r[mv].replace(",",""))))
PY
