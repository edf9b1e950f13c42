#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (both arms), ncu launch list + full captures.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_check.sh [tag]'
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total --format=csv | tail -1
nproc
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/pytest_gpu_$TAG.log
echo "== smoke"; timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -3 | tee $OUT/smoke_$TAG.log
echo "== bench reference"; timeout 400 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_ref_$TAG.json
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | tee $OUT/bench_$TAG.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_launches_$TAG.log 2>&1
echo "== ncu full (viterbi, lattice)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'kp_viterbi|kp_lattice_walk|kp_bucketize' -s 12 -c 4 \
    -o $OUT/prof_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu > $OUT/ncu_full_$TAG.log 2>&1
ls -la $OUT
