#!/bin/bash
# Round-end evidence: parity tests, sanitizer, bench both arms, ncu launch list + full capture of the
# dominant kernels.  usage: gpurun --timeout 1800 -- 'bash tools/gpu_final.sh r01c'
TAG=${1:-final}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/pytest_gpu_$TAG.log
echo "== smoke"; timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2 | tee $OUT/smoke_$TAG.log
echo "== compute-sanitizer memcheck (both device paths, 40 + 110 sentences + edge cases)"
timeout 900 compute-sanitizer --tool memcheck --target-processes all --launch-timeout 120 python tools/small_case.py 40 > $OUT/sanitize_$TAG.log 2>&1
grep -E "ERROR SUMMARY|PARITY|Error" $OUT/sanitize_$TAG.log | head -5
echo "== bench reference"; timeout 400 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 > $OUT/bench_ref_$TAG.json; cut -c1-400 $OUT/bench_ref_$TAG.json
echo "== bench"; timeout 600 python bench.py --steps 30 --warmup 3 2>&1 | tail -1 > $OUT/bench_$TAG.json; cut -c1-300 $OUT/bench_$TAG.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-parity > $OUT/ncu_launches_$TAG.log 2>&1
echo "== ncu full (viterbi, lattice fill)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'kp_viterbi|kp_lattice_walk|kp_lattice_count|kp_bucketize|kp_backtrace_find' -s 15 -c 5 \
    -o $OUT/prof_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu --no-parity > $OUT/ncu_full_$TAG.log 2>&1
ls -la $OUT | tail -12
