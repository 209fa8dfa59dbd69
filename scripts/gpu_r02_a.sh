#!/bin/bash
# Round 2, GPU call A: parity tests, the new default bench line, launch list and source-level ncu captures of the two
# kernels this round works on (composed_cfgmajor_kernel on C4, mesh_query_kernel on mesh10k).
set -u
OUT=gpurun_out/r02a
mkdir -p "$OUT"
timeout 1500 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?" >> "$OUT/pytest_gpu.log"
tail -5 "$OUT/pytest_gpu.log"
timeout 900 python bench.py --steps 20 --warmup 5 2> "$OUT/bench_default.err" | grep '^{' | tail -1 > "$OUT/bench_default_1gpu.jsonl"
head -c 600 "$OUT/bench_default_1gpu.jsonl"; echo
NCU="ncu --set full --clock-control none --import-source on"
timeout 500 $NCU -k regex:composed_cfgmajor -s 3 -c 1 -o "$OUT/c4_cfgmajor_v1" -f \
    python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > "$OUT/ncu_c4.log" 2>&1
timeout 500 $NCU -k regex:mesh_query_kernel -s 3 -c 1 -o "$OUT/mesh10k_v1" -f \
    python bench.py --workload mesh10k --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > "$OUT/ncu_mesh.log" 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$OUT/launches_bench_default.csv" \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
cp pytorch_volumetric_b200/csrc/libpvb.so "$OUT/libpvb_r02a.so"
ls -la "$OUT"
