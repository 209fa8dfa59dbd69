import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
torch.cuda.set_device(0)
wl = bench.make_workload("c4", 0, 1)
out = []
for s in wl.robot.sdf.sdfs:
    d = s._desc
    out.append({"flags": d.flags, "margin": d.prune_margin, "dims": list(d.dims), "bb_min": list(d.bb_min), "bb_max": list(d.bb_max),
                "valmin": float(s.voxels.raw_data.min()), "valmax": float(s.voxels.raw_data.max())})
v, g, w = wl.robot.sdf.query(wl.dev[0], cfg_begin=0, cfg_count=4, return_which=True)
out.append({"which_hist": torch.bincount(w, minlength=8).tolist(), "val_mean": float(v.mean())})
print(json.dumps(out))
