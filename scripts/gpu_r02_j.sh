#!/bin/bash
# Round 2, GPU call J: Morton resolution of the query binning (PVB_SORT_BITS) on the tree-walk workloads.
set -u
OUT=gpurun_out/r02j
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_gpu_mesh.py tests/test_gpu_chamfer_sample.py tests/test_gpu_baseline_parity.py -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?" >> "$OUT/pytest_gpu.log"
tail -3 "$OUT/pytest_gpu.log"
T="$OUT/tune_sort_bits.jsonl"; : > "$T"
for b in 6 7 8; do for w in mesh10k c5 mesh50k; do
  PVB_SORT_BITS=$b timeout 300 python scripts/tune_kernel.py $w 12 2>>"$OUT/tune.err" | grep '^{' | sed "s/^{/{\"sort_bits\": $b, /" >> "$T"
done; done
python - <<'PY'
import json
for l in open("gpurun_out/r02j/tune_sort_bits.jsonl"):
    d = json.loads(l); print(d["sort_bits"], d["workload"], "ms", round(d["ms_median"], 3))
PY
tail -3 "$OUT/tune.err"
