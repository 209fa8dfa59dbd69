#!/bin/bash
# Round 2, GPU call E: robot_serial_kernel (point-serial, nearest-sphere-first) -- parity + variants + ncu.
set -u
OUT=gpurun_out/r02e
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_gpu_composed.py tests/test_gpu_peer.py tests/test_gpu_baseline_parity.py tests/test_gpu_edge.py -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?" >> "$OUT/pytest_gpu.log"
tail -8 "$OUT/pytest_gpu.log"
T="$OUT/tune_c4.jsonl"; : > "$T"
run() { timeout 300 env "$@" python scripts/tune_c4_shapes.py 30 2>>"$OUT/tune.err" | grep '^{' >> "$T"; }
run PVB_ROBOT_KERNEL=0
run PVB_ROBOT_KERNEL=1 PVB_ROBOT_MIN_FILL=0
run PVB_ROBOT_KERNEL=2 PVB_ROBOT_MIN_FILL=0
for v in s1m4 s1m3 s0m3 s1m5; do
  run PVB_LIB=$PWD/tune/libpvb_rs_$v.so PVB_ROBOT_MIN_FILL=0
done
run PVB_ROBOT_KERNEL=2 PVB_ROBOT_MIN_FILL=0 PVB_ROBOT_WAVES=4
run PVB_ROBOT_KERNEL=2 PVB_ROBOT_MIN_FILL=0 PVB_ROBOT_WAVES=16
run PVB_ROBOT_KERNEL=2 PVB_ROBOT_MIN_FILL=0 PVB_ROBOT_WAVES=64
cat "$T"; tail -3 "$OUT/tune.err"
NCU="ncu --set full --clock-control none --import-source on"
timeout 500 $NCU -k regex:robot_serial -s 3 -c 1 -o "$OUT/c4_robot_serial_s1m4" -f \
    python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > "$OUT/ncu_c4.log" 2>&1
cp pytorch_volumetric_b200/csrc/libpvb.so "$OUT/libpvb_r02e.so"
ls -la "$OUT"
