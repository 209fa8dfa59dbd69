#!/bin/bash
# Round 2, GPU call O: grid width (PVB_ROBOT_WAVES) and step size (PVB_ROBOT_FINE_STEPS) of robot_serial_kernel on the
# per-rank slabs (200 / 100 / 50 / 25 configurations), kernel-only times.
set -u
OUT=gpurun_out/r02o
mkdir -p "$OUT"
T="$OUT/tune_c4_slabs.jsonl"; : > "$T"
for f in 12 0 100000; do for w in 1 2 3 4 6 8; do
  PVB_ROBOT_WAVES=$w PVB_ROBOT_FINE_STEPS=$f timeout 200 python scripts/tune_c4_slabs.py 15 2>>"$OUT/tune.err" | grep '^{' >> "$T"
done; done
cat "$T"
tail -3 "$OUT/tune.err"
