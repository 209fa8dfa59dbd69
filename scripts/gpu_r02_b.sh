#!/bin/bash
# Round 2, GPU call B: new robot_query_kernel + pvb_fk_serial -- parity tests, variant timing, ncu, default bench.
set -u
OUT=gpurun_out/r02b
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_gpu_composed.py tests/test_gpu_peer.py tests/test_gpu_baseline_parity.py tests/test_gpu_edge.py -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?" >> "$OUT/pytest_gpu.log"
tail -15 "$OUT/pytest_gpu.log"
T="$OUT/tune_c4.jsonl"; : > "$T"
run() { timeout 300 env "$@" python scripts/tune_c4_shapes.py 30 2>>"$OUT/tune.err" | grep '^{' >> "$T"; }
run PVB_ROBOT_KERNEL=0
run PVB_ROBOT_KERNEL=1
run PVB_ROBOT_KERNEL=1 PVB_ROBOT_MIN_FILL=0
run PVB_ROBOT_KERNEL=1 PVB_ROBOT_MIN_FILL=0 PVB_ROBOT_WAVES=2
run PVB_ROBOT_KERNEL=1 PVB_ROBOT_MIN_FILL=0 PVB_ROBOT_WAVES=8
run PVB_ROBOT_KERNEL=1 PVB_ROBOT_MIN_FILL=0 PVB_ROBOT_WAVES=16
for v in p4m3 p2m4 p2m5 p2m6; do
  run PVB_LIB=$PWD/tune/libpvb_rb_$v.so PVB_ROBOT_MIN_FILL=0
done
cat "$T"
NCU="ncu --set full --clock-control none --import-source on"
timeout 500 $NCU -k regex:robot_query -s 3 -c 1 -o "$OUT/c4_robot_p4m4" -f \
    python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > "$OUT/ncu_c4.log" 2>&1
PVB_LIB=$PWD/tune/libpvb_rb_p2m5.so timeout 500 $NCU -k regex:robot_query -s 3 -c 1 -o "$OUT/c4_robot_p2m5" -f \
    python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > "$OUT/ncu_c4b.log" 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2> "$OUT/bench_default.err" | grep '^{' | tail -1 > "$OUT/bench_default_1gpu.jsonl"
head -c 400 "$OUT/bench_default_1gpu.jsonl"; echo
cp pytorch_volumetric_b200/csrc/libpvb.so "$OUT/libpvb_r02b.so"
ls -la "$OUT"
