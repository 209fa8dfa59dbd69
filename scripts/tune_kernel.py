"""Times one workload's dominant kernel for the libpvb.so named by PVB_LIB (tuning only).
usage: PVB_LIB=tune/libpvb_X.so python scripts/tune_kernel.py c2|mesh10k|c4|c5|c3 [iters]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
torch.cuda.set_device(0)
wl = bench.make_workload(name, 0, 1)
for i in range(5):
    wl.step(i)
torch.cuda.synchronize()
ts = []
for i in range(iters):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    wl.step(i)
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ts.sort()
med = ts[len(ts) // 2]
print(json.dumps({"lib": os.environ.get("PVB_LIB", "default"), "stage": os.environ.get("PVB_STAGE_NODES"),
                  "workload": name, "ms_median": med, "ms_min": ts[0],
                  "GBps": wl.alg_bytes / med / 1e6, "units_per_s": wl.units / med * 1e3}))
