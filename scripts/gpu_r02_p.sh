#!/bin/bash
# Round 2, GPU call P: host-side pipelines (ComposedSDF result streamed out slab by slab; MeshSDF host batches in
# chunks; shared _HostChunkStream): full GPU test tier + the bench line's e2e values.
set -u
OUT=gpurun_out/r02p
mkdir -p "$OUT"
timeout 1200 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?" >> "$OUT/pytest_gpu.log"
tail -5 "$OUT/pytest_gpu.log"
timeout 900 python bench.py --no-cpu-baseline > "$OUT/bench_default_1gpu.jsonl" 2> "$OUT/bench.err"
python - <<'PY'
import json
for l in open("gpurun_out/r02p/bench_default_1gpu.jsonl"):
    if not l.startswith("{"): continue
    d = json.loads(l)
    print("c4 ms", round(d["ms_per_step"], 4), "kernel_ms", d["roofline"].get("kernel_ms"), "e2e ms", round(d["e2e"]["ms_per_step"], 3))
    for k, v in d["workloads"].items():
        print(k, "ms", round(v["ms_per_step"], 4), "kernel_ms", v["roofline"].get("kernel_ms"), "e2e ms", round(v["e2e"]["ms_per_step"], 3))
PY
tail -3 "$OUT/bench.err"
