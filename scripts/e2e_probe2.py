import os, sys, time, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
torch.cuda.set_device(0)
wl = bench.make_workload("c2", 0, 1)
h = wl.host[0]
res = {}
def t(fn, reps=6):
    for _ in range(3):
        r = fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return [round(x, 2) for x in ts]
for chunk in (1 << 21, 1 << 19, 1 << 20, 3 << 19, 1 << 21, 1 << 20):
    wl.sdf_host.host_pipeline_chunk = chunk
    if hasattr(wl.sdf_host, "_pipe_state"): del wl.sdf_host._pipe_state
    res[f"chunk{chunk}_{len(res)}"] = t(lambda: wl.sdf_host(h))
print(json.dumps(res))
