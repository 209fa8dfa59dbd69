#!/bin/bash
# Round 2, GPU call K: instruction diet of robot_serial_kernel (packed nearest-sphere key, no per-link `on` / n_sdf
# tests, division-free flush index, staging pointers): parity, C4 time, instruction count.
set -u
OUT=gpurun_out/r02k
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_gpu_composed.py tests/test_gpu_baseline_parity.py tests/test_gpu_edge.py -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?" >> "$OUT/pytest_gpu.log"
tail -3 "$OUT/pytest_gpu.log"
T="$OUT/tune_c4_diet.jsonl"; : > "$T"
for rep in 1 2 3; do
  PVB_LIB=tune/libpvb_base.so timeout 300 python scripts/tune_kernel.py c4 20 2>>"$OUT/tune.err" | grep '^{' >> "$T"
  timeout 300 python scripts/tune_kernel.py c4 20 2>>"$OUT/tune.err" | grep '^{' >> "$T"
done
python - <<'PY'
import json
for l in open("gpurun_out/r02k/tune_c4_diet.jsonl"):
    d = json.loads(l); print(d["lib"], d["workload"], "ms", round(d["ms_median"], 4), "min", round(d.get("ms_min", 0), 4))
PY
timeout 600 ncu --set full --import-source on --clock-control none -k regex:robot_serial -c 1 -o "$OUT/robot_serial_c4_diet" python scripts/tune_kernel.py c4 2 > "$OUT/ncu.log" 2>&1
ncu -i "$OUT/robot_serial_c4_diet.ncu-rep" --page raw --csv > "$OUT/robot_serial_c4_diet.raw.csv" 2>/dev/null
python scripts/ncu_lines.py "$OUT/robot_serial_c4_diet.raw.csv" robot_serial 2>/dev/null | tail -40
tail -3 "$OUT/tune.err"
