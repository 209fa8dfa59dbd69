#!/bin/bash
# Round 2, multi-GPU call (N = number of GPUs of the box): peer / symmetric-memory / multicast re-assembly -- bit-exact
# check and timing at the C4 shape, then the default bench line.
set -u
N=${1:-2}
OUT=gpurun_out/r02d_n$N
mkdir -p "$OUT"
export NCCL_DEBUG=WARN
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_gpu_composed.py -m gpu -x -q > "$OUT/pytest_composed.log" 2>&1; echo "pytest rc=$?" >> "$OUT/pytest_composed.log"
  tail -3 "$OUT/pytest_composed.log"


  timeout 600 python -m pytest tests/test_gpu_peer.py -m gpu -x -q > "$OUT/pytest_peer.log" 2>&1; echo "pytest rc=$?" >> "$OUT/pytest_peer.log"
  tail -5 "$OUT/pytest_peer.log"
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641 \
    scripts/peer_check.py 200 100000 20 > "$OUT/peer_check_c4.log" 2>&1; echo "peer_check rc=$?" >> "$OUT/peer_check_c4.log"
grep -E "^\{|PEER_OK|rc=|Error|error" "$OUT/peer_check_c4.log" | cut -c1-1500
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29642 \
    bench.py --gpus $N --steps 20 --warmup 5 2> "$OUT/bench.err" | grep '^{' | tail -1 > "$OUT/bench_default_${N}gpu.jsonl"
python - <<PY
import json
d = json.loads(open("$OUT/bench_default_${N}gpu.jsonl").read())
print("value", d["value"], "ms", d["ms_per_step"], "chosen", d["reassembly"]["chosen"])
for k, v in d["reassembly"].items():
    print(" ", k, v)
print("e2e", d["e2e"])
for k, w in (d.get("workloads") or {}).items():
    print(" ", k, {kk: w.get(kk) for kk in ("ms_per_step", "value")}, "e2e", (w.get("e2e") or {}).get("ms_per_step"))
PY
tail -5 "$OUT/bench.err"
ls -la "$OUT"
