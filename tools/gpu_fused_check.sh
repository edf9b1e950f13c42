#!/bin/bash
OUT=gpurun_out
echo "== fused tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "both_paths or fused_path or cfg1 or ragged or full_cfg2 or long_unknown or first_character or fuzz or compact or reference_fixture" 2>&1 | tail -5
echo "== device sweep"; timeout 600 python tools/device_sweep.py 2>&1 | tail -9
echo "== latency sweep"; timeout 900 python tools/latency_sweep.py > $OUT/latency_sweep_x.json 2> $OUT/latency_sweep_x.err; head -5 $OUT/latency_sweep_x.err | cut -c1-260
