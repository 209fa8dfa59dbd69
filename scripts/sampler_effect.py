"""Does clock sampling perturb the timed loop?  (tuning aid)"""
import os, sys, time, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
torch.cuda.set_device(0)
wl = bench.make_workload("c2", 0, 1)
for _ in range(10): wl.step(0)

def run(steps, period):
    s = bench.ClockSampler(0, period_s=period) if period else None
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if s: s.start()
    a.record()
    for i in range(steps):
        out = wl.step(i)
    b.record()
    torch.cuda.synchronize()
    c = s.stop() if s else None
    return a.elapsed_time(b) / steps * 1e3, c

res = {}
for period in (None, 0.002, 0.02, 0.1, None):
    for rep in range(2):
        us, c = run(200, period)
        res[f"period={period} rep{rep}"] = (round(us, 2), c)
print(json.dumps(res))
