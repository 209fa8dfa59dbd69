#!/bin/bash
# Round 2, GPU call R: robot_serial_kernel with 4-point staging tiles (35.8 KB of shared memory per block instead of
# 52 KB) at 4 / 5 / 6 blocks per SM -- the first time a fifth block actually fits (the earlier 48-register build was
# still limited to 4 blocks by its 52 KB of shared memory).
set -u
OUT=gpurun_out/r02r
mkdir -p "$OUT"
T="$OUT/tune_c4_occupancy.jsonl"; : > "$T"
for v in default c4m4 c4m5 c4m6; do
  if [ "$v" = default ]; then L=""; else L="tune/libpvb_$v.so"; fi
  PVB_LIB=$L timeout 200 python scripts/tune_c4_slabs.py 15 2>>"$OUT/tune.err" | grep '^{' | sed "s/^{/{\"variant\": \"$v\", /" >> "$T"
done
cat "$T"
tail -3 "$OUT/tune.err"
