"""After `scripts/profile_round.sh <round>` ran on the GPU box: summarise every .ncu-rep of gpurun_out/<round>/ into
profiles/<round>/*.ncu.json, copy the bench lines / launch list, and refresh profiles/ncu_traffic.json (DRAM bytes per
launch of each workload's dominant kernel, read by bench.py into roofline.traffic) and profiles/ncu_limits.json (issue
slots / L1 data pipe / DRAM in % of peak from the same capture, read into roofline.ncu).
usage: python scripts/collect_profiles.py r02"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rnd = sys.argv[1] if len(sys.argv) > 1 else "r02"
src = os.path.join(ROOT, "gpurun_out", rnd)
dst = os.path.join(ROOT, "profiles", rnd)
os.makedirs(dst, exist_ok=True)
WORKLOAD_OF = {"robot_serial_c4": "c4", "grid_lookup_tma_c2": "c2", "composed_query_c3": "c3",
               "composed_query_c3cached": "c3cached", "mesh_query_mesh10k": "mesh10k", "chamfer_partial_c5": "c5"}
traffic_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
# what actually bounds each kernel (the tree walks and the RobotSDF kernel are not HBM-bound): % of peak of the issue
# slots, the L1 data pipe and DRAM from the same capture; bench.py copies them into roofline.ncu
limits_path = os.path.join(ROOT, "profiles", "ncu_limits.json")
limits = json.load(open(limits_path)) if os.path.exists(limits_path) else {}
LIMIT_KEYS = {"issue_slots_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
              "l1_data_pipe_pct": "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
              "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
              "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active"}
for f in sorted(os.listdir(src)):
    p = os.path.join(src, f)
    if f.endswith(".raw.csv"):          # raw metric page written on the box (the report itself may not have travelled)
        name = f[:-8]
        rows = list(csv.reader(io.StringIO(open(p).read())))
        if len(rows) > 2 and name in WORKLOAD_OF:
            idx = {h: i for i, h in enumerate(rows[0])}
            unit = {h: u for h, u in zip(rows[0], rows[1])}
            tot = 0.0
            for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                v = float(rows[2][idx[k]])
                tot += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit[k]]
            traffic[WORKLOAD_OF[name]] = int(tot)
            lim = {k: round(float(rows[2][idx[m]]), 1) for k, m in LIMIT_KEYS.items() if m in idx}
            lim["source"] = f"profiles/{rnd}/{name}.ncu.json"
            limits[WORKLOAD_OF[name]] = lim
        print("summarised", f)
    elif f.endswith((".jsonl", ".csv", ".ncu.json")) or f.startswith("pytest"):
        shutil.copy(p, os.path.join(dst, f))
json.dump(traffic, open(traffic_path, "w"), indent=1, sort_keys=True)
json.dump(limits, open(limits_path, "w"), indent=1, sort_keys=True)
print("ncu_traffic.json:", traffic)
