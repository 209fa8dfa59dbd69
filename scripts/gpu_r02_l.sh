#!/bin/bash
# Round 2, GPU call L: robot_serial_kernel variants after the instruction diet -- gradient rotated inside the visit
# (no re-read of the winner's transform rows), the same with 3 blocks per SM (80 registers).
set -u
OUT=gpurun_out/r02l
mkdir -p "$OUT"
T="$OUT/tune_c4_rot.jsonl"; : > "$T"
for rep in 1 2 3; do for v in diet rot rot3; do
  PVB_LIB=tune/libpvb_$v.so timeout 300 python scripts/tune_kernel.py c4 20 2>>"$OUT/tune.err" | grep '^{' >> "$T"
done; done
python - <<'PY'
import json
for l in open("gpurun_out/r02l/tune_c4_rot.jsonl"):
    d = json.loads(l); print(d["lib"], d["workload"], "ms", round(d["ms_median"], 4), "min", round(d.get("ms_min", 0), 4))
PY
tail -3 "$OUT/tune.err"
