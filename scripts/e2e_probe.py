import os, sys, time, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
torch.cuda.set_device(0)
wl = bench.make_workload("c2", 0, 1)
res = {}
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
h = wl.host[0]
res["pinned_alloc_160MB_ms"] = t(lambda: (torch.empty(10_000_000, pin_memory=True), torch.empty(10_000_000, 3, pin_memory=True)))
d = torch.empty_like(wl.dev[0])
res["h2d_120MB_ms"] = t(lambda: d.copy_(h, non_blocking=True))
outp = torch.empty(10_000_000, 3, pin_memory=True)
res["d2h_120MB_ms"] = t(lambda: outp.copy_(d, non_blocking=True))
for chunk in (1 << 20, 1 << 21, 1 << 22):
    wl.sdf_host.host_pipeline_chunk = chunk
    if hasattr(wl.sdf_host, "_pipe_state"): del wl.sdf_host._pipe_state
    res[f"pipeline_chunk{chunk}_ms"] = t(lambda: wl.sdf_host(h))
wl.sdf_host.host_pipeline_min_points = 1 << 40
res["plain_ms"] = t(lambda: wl.sdf_host(h))
# hold previous result like the bench does
def held():
    global keep
    keep = wl.sdf_host(h)
res["plain_held_ms"] = t(held)
print(json.dumps(res))
