#!/bin/bash
# Round 2, GPU call C2: robot_query_kernel (direct stores, rolled link loop) variants.
set -u
OUT=gpurun_out/r02c
mkdir -p "$OUT"
PVB_LIB=$PWD/tune/libpvb_rb_r_p2m4.so timeout 600 python -m pytest tests/test_gpu_composed.py tests/test_gpu_peer.py -m gpu -x -q > "$OUT/pytest_gpu_p2.log" 2>&1; echo "pytest rc=$?" >> "$OUT/pytest_gpu_p2.log"
tail -3 "$OUT/pytest_gpu_p2.log"
T="$OUT/tune_c4.jsonl"; : > "$T"
run() { timeout 300 env "$@" python scripts/tune_c4_shapes.py 30 2>>"$OUT/tune.err" | grep '^{' >> "$T"; }
run PVB_ROBOT_KERNEL=0
for v in r_p4m4 r_p4m3 r_p2m4 r_p2m5 u_p2m4; do
  run PVB_LIB=$PWD/tune/libpvb_rb_$v.so PVB_ROBOT_MIN_FILL=0
done
run PVB_LIB=$PWD/tune/libpvb_rb_r_p2m4.so PVB_ROBOT_MIN_FILL=0 PVB_ROBOT_WAVES=8
run PVB_LIB=$PWD/tune/libpvb_rb_r_p4m3.so PVB_ROBOT_MIN_FILL=0 PVB_ROBOT_WAVES=8
run PVB_LIB=$PWD/tune/libpvb_rb_r_p2m4.so PVB_ROBOT_MIN_FILL=0 PVB_ROBOT_WAVES=2
cat "$T"
NCU="ncu --set full --clock-control none --import-source on"
PVB_LIB=$PWD/tune/libpvb_rb_r_p2m4.so timeout 500 $NCU -k regex:robot_query -s 3 -c 1 -o "$OUT/c4_robot_r_p2m4" -f \
    python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > "$OUT/ncu_c4.log" 2>&1
PVB_LIB=$PWD/tune/libpvb_rb_r_p4m3.so timeout 500 $NCU -k regex:robot_query -s 3 -c 1 -o "$OUT/c4_robot_r_p4m3" -f \
    python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > "$OUT/ncu_c4b.log" 2>&1
ls -la "$OUT"
