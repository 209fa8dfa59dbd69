#!/bin/bash
# Round 2, GPU call M: full ncu capture (with source) of mesh_query_kernel on mesh10k, for per-line attribution.
set -u
OUT=gpurun_out/r02m
mkdir -p "$OUT"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:mesh_query -c 1 -o "$OUT/mesh_query_mesh10k" python scripts/tune_kernel.py mesh10k 2 > "$OUT/ncu.log" 2>&1
tail -3 "$OUT/ncu.log"
cp pytorch_volumetric_b200/csrc/libpvb.so "$OUT/libpvb_r02m.so"
ls -la "$OUT"
