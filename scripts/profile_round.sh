#!/bin/bash
# One GPU-box call that refreshes the judged evidence of a round.  Run from the repo root ON THE GPU BOX:
#
#   gpurun --timeout 1800 -- 'bash scripts/profile_round.sh r02'
#
# and afterwards, in the build container:  python scripts/collect_profiles.py r02
# Numbers printed by anything that ran under ncu are never bench values: the .jsonl files come from separate runs.
set -u
R=${1:-r02}
OUT=gpurun_out/$R
mkdir -p "$OUT"
NCU="ncu --set full --clock-control none --import-source on"

# 0. the GPU test tier
timeout 1500 python -m pytest tests -m gpu -q > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?" >> "$OUT/pytest_gpu.log"
tail -4 "$OUT/pytest_gpu.log"

# 1. the default bench line (not under a profiler): headline C4 + every other BASELINE config + CPU baselines
timeout 1200 python bench.py --steps 20 --warmup 5 2> "$OUT/bench_default.err" | grep '^{' | tail -1 > "$OUT/bench_default_1gpu.jsonl"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 2> "$OUT/bench_reference.err" | grep '^{' | tail -1 > "$OUT/bench_reference.jsonl"
timeout 300 python bench.py --workload c4readme --steps 50 --no-cpu-baseline 2>/dev/null | grep '^{' | tail -1 > "$OUT/bench_c4readme.jsonl"
timeout 300 python bench.py --workload mesh50k --steps 20 --no-cpu-baseline 2>/dev/null | grep '^{' | tail -1 > "$OUT/bench_mesh50k.jsonl"

# 2. launch list of the default bench command (per-launch device times, cold cache, serialised)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file "$OUT/launches_bench_default.csv" \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1

# 3. one full capture per dominant kernel (a single launch each: ncu replays it ~40 times)
# (gpurun brings back at most 64 MiB: every report is summarised on the box -- scripts/ncu_summary.py, plus the raw
#  metric page as csv -- and only the report named in KEEP_REP travels)
KEEP_REP=${KEEP_REP:-robot_serial_c4}
capture() {   # name, kernel regex, workload
    timeout 500 $NCU -k "regex:$2" -s 3 -c 1 -o "$OUT/$1" -f \
        python bench.py --workload "$3" --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > "$OUT/$1.log" 2>&1
    python scripts/ncu_summary.py "$OUT/$1.ncu-rep" "$OUT/$1.ncu.json" > /dev/null 2>&1
    ncu -i "$OUT/$1.ncu-rep" --page raw --csv > "$OUT/$1.raw.csv" 2>/dev/null
    [ "$1" = "$KEEP_REP" ] || rm -f "$OUT/$1.ncu-rep"
}
capture robot_serial_c4      robot_serial_kernel         c4
capture grid_lookup_tma_c2   grid_lookup_tma_kernel      c2
capture composed_query_c3    composed_query_kernel       c3
capture composed_query_c3cached composed_query_kernel    c3cached
capture mesh_query_mesh10k   mesh_query_kernel           mesh10k
capture chamfer_partial_c5   chamfer_partial_kernel      c5
[ -f "$OUT/$KEEP_REP.ncu-rep" ] && cp pytorch_volumetric_b200/csrc/libpvb.so "$OUT/libpvb_$R.so"
du -sh "$OUT"
ls -la "$OUT"
