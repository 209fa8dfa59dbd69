#!/bin/bash
# Round 2, GPU call N: triangles per BVH4 leaf (PVB_BVH_LEAF 1..4) on the tree-walk workloads.
set -u
OUT=gpurun_out/r02n
mkdir -p "$OUT"
T="$OUT/tune_bvh_leaf.jsonl"; : > "$T"
for b in 4 3 2 1; do for w in mesh10k c5 c3 mesh50k; do
  PVB_BVH_LEAF=$b timeout 300 python scripts/tune_kernel.py $w 12 2>>"$OUT/tune.err" | grep '^{' | sed "s/^{/{\"bvh_leaf\": $b, /" >> "$T"
done; done
python - <<'PY'
import json
for l in open("gpurun_out/r02n/tune_bvh_leaf.jsonl"):
    d = json.loads(l); print(d["bvh_leaf"], d["workload"], "ms", round(d["ms_median"], 3))
PY
tail -3 "$OUT/tune.err"
