"""Kernel-only time (library event pair, no host overhead) of RobotSDF queries of the C4 arm for the per-rank slabs of
1 / 2 / 4 / 8 GPUs.  usage: [PVB_ROBOT_WAVES=w] [PVB_ROBOT_FINE_STEPS=f] python scripts/tune_c4_slabs.py [iters]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pytorch_volumetric_b200 import _native as nat  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 15
torch.cuda.set_device(0)
wl = bench.make_workload("c4", 0, 1)
nat.timing_enable(True)
res = {}
for n in (200, 100, 50, 25):
    for i in range(4):
        wl.robot.sdf.query(wl.dev[i % 3], cfg_begin=0, cfg_count=n)
    ts = []
    for i in range(iters):
        wl.robot.sdf.query(wl.dev[i % 3], cfg_begin=0, cfg_count=n)
        ts.append(nat.timing_last_ms())
    ts.sort()
    res[n] = round(ts[len(ts) // 2], 4)
print(json.dumps({"waves": os.environ.get("PVB_ROBOT_WAVES"), "fine_steps": os.environ.get("PVB_ROBOT_FINE_STEPS"),
                  "kernel_ms_by_cfg_count": res}))
