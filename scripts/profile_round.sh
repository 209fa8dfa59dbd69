#!/bin/bash
# One GPU-box call that refreshes the judged evidence of a round.  Run from the repo root ON THE GPU BOX:
#
#   gpurun --timeout 900 -- 'bash scripts/profile_round.sh r02'
#
# and afterwards, in the build container:
#
#   for f in gpurun_out/r02/*.ncu-rep; do python scripts/ncu_summary.py $f profiles/r02/$(basename ${f%.ncu-rep}).ncu.json; done
#   cp gpurun_out/r02/*.jsonl gpurun_out/r02/*.csv profiles/r02/
#
# Numbers printed by anything that ran under ncu are never bench values: the .jsonl files come from separate runs.
set -u
R=${1:-r02}
OUT=gpurun_out/$R
mkdir -p "$OUT"
NCU="ncu --set full --clock-control none --import-source on"

# 1. bench lines (not under a profiler)
python bench.py                         2>/dev/null | grep '^{' | tail -1 > "$OUT/bench_default_1gpu.jsonl"
for w in mesh10k c3 c3cached c4 c4readme c5; do
    timeout 300 python bench.py --workload $w --steps 30 --no-cpu-baseline 2>/dev/null | grep '^{' | tail -1
done > "$OUT/bench_workloads_1gpu.jsonl"

# 2. launch list of the default bench command (per-launch device times, cold cache, serialised)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches_bench_c2.csv" \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1

# 3. one full capture per dominant kernel (a single launch each: ncu replays it ~40 times)
capture() {   # name, kernel regex, workload
    timeout 600 $NCU -k "regex:$2" -s 3 -c 1 -o "$OUT/$1" -f \
        python bench.py --workload "$3" --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > "$OUT/$1.log" 2>&1
}
capture grid_lookup_tma      grid_lookup_tma_kernel      c2
capture composed_cfgmajor_c4 composed_cfgmajor_kernel    c4
capture composed_query_c3    composed_query_kernel       c3
capture mesh_query           mesh_query_kernel           mesh10k
capture chamfer_partial_c5   chamfer_partial_kernel      c5
ls -la "$OUT"
