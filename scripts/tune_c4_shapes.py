"""Times RobotSDF queries of the C4 arm for several configuration counts (the per-rank slabs of 1 / 2 / 4 / 8 GPUs).
usage: [PVB_LIB=...] [PVB_ROBOT_KERNEL=0|1] [PVB_ROBOT_MIN_FILL=x] python scripts/tune_c4_shapes.py [iters]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 30
torch.cuda.set_device(0)
wl = bench.make_workload("c4", 0, 1)
res = {}
for n in (200, 100, 50, 25, 8, 3):
    fn = lambda i: wl.robot.sdf.query(wl.dev[i % 3], cfg_begin=0, cfg_count=n)      # noqa: E731
    for i in range(5):
        fn(i)
    torch.cuda.synchronize()
    ts = []
    for i in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(i); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    res[n] = round(ts[len(ts) // 2], 4)
print(json.dumps({"lib": os.path.basename(os.environ.get("PVB_LIB", "default")),
                  "robot_kernel": os.environ.get("PVB_ROBOT_KERNEL", "1"), "min_fill": os.environ.get("PVB_ROBOT_MIN_FILL"),
                  "waves": os.environ.get("PVB_ROBOT_WAVES"), "min_cfg": os.environ.get("PVB_ROBOT_MIN_CFG"),
                  "ms_by_cfg_count": res}))
