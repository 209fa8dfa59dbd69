import time, json, torch
torch.cuda.set_device(0)
d = torch.empty(10_000_000, 3, device="cuda")
s2 = torch.cuda.Stream()
def it(side, hold):
    global kept
    t0 = time.perf_counter()
    a = torch.empty(10_000_000, 3, pin_memory=True)
    t1 = time.perf_counter()
    if side:
        s2.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s2):
            a[:5_000_000].copy_(d[:5_000_000], non_blocking=True)
            a[5_000_000:].copy_(d[5_000_000:], non_blocking=True)
        s2.synchronize()
    else:
        a.copy_(d, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    t2 = time.perf_counter()
    if hold: kept = a
    return round((t1 - t0) * 1e3, 2), round((t2 - t1) * 1e3, 2)
res = {}
for side in (False, True):
    for hold in (False, True):
        res[f"side={side} hold={hold}"] = [it(side, hold) for _ in range(6)]
print(json.dumps(res))
