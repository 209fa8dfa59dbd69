"""Host-side overhead of the public API per call (tuning aid): wall time per call on a tiny input, and
back-to-back device time per step on the C2 batch with / without per-step events."""
import os, sys, time, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

torch.cuda.set_device(0)
wl = bench.make_workload("c2", 0, 1)
tiny = wl.dev[0][:4096].contiguous()
for _ in range(20):
    wl.sdf(tiny)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(2000):
    wl.sdf(tiny)
torch.cuda.synchronize()
per_call_us = (time.perf_counter() - t0) / 2000 * 1e6

def run(steps, per_step_events):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    a.record()
    for i in range(steps):
        if per_step_events: evs[i][0].record()
        out = wl.step(i)
        if per_step_events: evs[i][1].record()
    b.record()
    t_cpu = time.perf_counter() - t0
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps * 1e3, t_cpu / steps * 1e6

for _ in range(5): wl.step(0)
res = {"tiny_call_wall_us": per_call_us}
for pse in (False, True):
    for rep in range(3):
        gpu_us, cpu_us = run(30, pse)
        res[f"b2b_events={pse}_rep{rep}"] = {"gpu_us_per_step": gpu_us, "cpu_issue_us_per_step": cpu_us}
print(json.dumps(res, indent=1))
