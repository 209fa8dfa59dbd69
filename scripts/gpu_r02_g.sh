#!/bin/bash
# Round 2, GPU call G (1 GPU): fine steps for small launches -- composed/peer parity in both step sizes, shapes timing.
set -u
OUT=gpurun_out/r02g
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_gpu_composed.py tests/test_gpu_peer.py tests/test_gpu_baseline_parity.py -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?" >> "$OUT/pytest_gpu.log"
tail -4 "$OUT/pytest_gpu.log"
PVB_ROBOT_FINE_STEPS=0 timeout 900 python -m pytest tests/test_gpu_composed.py tests/test_gpu_peer.py -m gpu -x -q > "$OUT/pytest_gpu_coarse.log" 2>&1; echo "pytest rc=$?" >> "$OUT/pytest_gpu_coarse.log"
tail -3 "$OUT/pytest_gpu_coarse.log"
PVB_ROBOT_FINE_STEPS=1000000 timeout 900 python -m pytest tests/test_gpu_composed.py -m gpu -x -q > "$OUT/pytest_gpu_fine.log" 2>&1; echo "pytest rc=$?" >> "$OUT/pytest_gpu_fine.log"
tail -3 "$OUT/pytest_gpu_fine.log"
T="$OUT/tune_c4.jsonl"; : > "$T"
run() { timeout 300 env "$@" python scripts/tune_c4_shapes.py 30 2>>"$OUT/tune.err" | grep '^{' >> "$T"; }
run PVB_ROBOT_FINE_STEPS=0
run PVB_ROBOT_FINE_STEPS=6
run PVB_ROBOT_FINE_STEPS=12
run PVB_ROBOT_FINE_STEPS=1000000
cat "$T"; tail -3 "$OUT/tune.err"
ls -la "$OUT"
