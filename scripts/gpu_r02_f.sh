#!/bin/bash
# Round 2, GPU call F: robot_serial_kernel with lane-split remainder tiles -- full GPU suite, shapes, ncu, default bench.
set -u
OUT=gpurun_out/r02f
mkdir -p "$OUT"
timeout 1200 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?" >> "$OUT/pytest_gpu.log"
tail -8 "$OUT/pytest_gpu.log"
T="$OUT/tune_c4.jsonl"; : > "$T"
run() { timeout 300 env "$@" python scripts/tune_c4_shapes.py 30 2>>"$OUT/tune.err" | grep '^{' >> "$T"; }
run PVB_ROBOT_KERNEL=0
run PVB_ROBOT_KERNEL=2
run PVB_ROBOT_KERNEL=2 PVB_ROBOT_MIN_CFG=1
run PVB_ROBOT_KERNEL=2 PVB_ROBOT_WAVES=8
run PVB_ROBOT_KERNEL=2 PVB_ROBOT_WAVES=2
cat "$T"; tail -3 "$OUT/tune.err"
NCU="ncu --set full --clock-control none --import-source on"
timeout 500 $NCU -k regex:robot_serial -s 3 -c 1 -o "$OUT/c4_robot_serial_lc" -f \
    python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > "$OUT/ncu_c4.log" 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 2> "$OUT/bench_default.err" | grep '^{' | tail -1 > "$OUT/bench_default_1gpu.jsonl"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02f/bench_default_1gpu.jsonl").read())
print("headline ms", d["ms_per_step"], "kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "e2e ms", d["e2e"]["ms_per_step"], "reconf", d["reconfigure_and_query"]["ms_per_step"])
for k, w in d["workloads"].items():
    print(k, "ms", round(w["ms_per_step"], 4), "kernel", round(w["roofline"]["kernel_ms"], 4), "frac", round(w["roofline"]["frac"], 4), "e2e", round(w["e2e"]["ms_per_step"], 3))
PY
cp pytorch_volumetric_b200/csrc/libpvb.so "$OUT/libpvb_r02f.so"
ls -la "$OUT"
