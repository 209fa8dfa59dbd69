#!/bin/bash
# Round 2, GPU call I: nearest-bound-first evaluation of mesh sub-SDFs in composed_query_kernel<true> (C3).
set -u
OUT=gpurun_out/r02i
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_gpu_composed.py tests/test_gpu_baseline_parity.py tests/test_gpu_edge.py -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?" >> "$OUT/pytest_gpu.log"
tail -4 "$OUT/pytest_gpu.log"
for nf in 0 1; do PVB_COMP_NEAREST_FIRST=$nf timeout 300 python bench.py --workload c3 --steps 20 --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('nearest_first=$nf', 'ms', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms'],3))"; done | tee "$OUT/c3_nearest_first.log"
