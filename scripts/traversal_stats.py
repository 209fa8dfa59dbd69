"""Node visits and triangle tests per query of the tree-walk device code, counted by running the
device header on the host (tests/hostsim, PVB_STAT hooks): the same bvh_closest / bvh_parity_x / bvh_parity the
kernels execute, on the bench workloads' meshes and query distributions.  CPU only."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import workloads  # noqa: E402
import hostsim_lib as hs  # noqa: E402
import pytorch_volumetric_b200 as pv  # noqa: E402
from pytorch_volumetric_b200 import _native as nat  # noqa: E402


def run(label, v, f, pts, mode):
    obj = pv.MeshObjectFactory(label, mesh=(v, f))
    d, keep = hs.mesh_desc(obj)
    *_, st = hs.mesh_query(d, pts, mode=mode, stats=True)
    nodes, tris, depth = obj._bvh_host
    out = {"workload": label, "triangles": int(len(f)), "bvh4_nodes": int(len(nodes)), "bvh4_depth": int(depth),
           "closed": bool(obj.is_closed), "queries": int(len(pts)), **{k: round(x, 2) for k, x in st.items()}}
    out["bytes_touched_per_query"] = round(128 * (st["closest_nodes"] + st["parity_nodes"]) +
                                           48 * (st["closest_tris"] + st["parity_tris"]))
    print(json.dumps(out))


def seeded(label, v, f, pts):
    """The irreducible part of the closest-point walk: the same traversal started with the TRUE squared distance (of a
    first, unseeded pass; 1 ulp of slack) as its search radius.  What is left are the leaves whose boxes reach inside
    the answer -- no seeding scheme, query ordering or first-hit heuristic can test fewer triangles than that."""
    import ctypes
    obj = pv.MeshObjectFactory(label, mesh=(v, f))
    d, keep = hs.mesh_desc(obj)
    dist, *_ = hs.mesh_query(d, pts, mode=0)
    p = np.ascontiguousarray(pts.numpy() if torch.is_tensor(pts) else pts, dtype=np.float32)
    out = {"workload": label, "queries": int(len(p))}
    for name, init in (("unseeded", None), ("seeded_with_answer", np.nextafter((dist.astype(np.float32)) ** 2, np.float32(np.inf)))):
        st = np.zeros(2, np.int64)
        hs.lib().sim_closest_seeded(ctypes.byref(d), p.ctypes.data_as(ctypes.c_void_p), ctypes.c_longlong(len(p)),
                                    init.ctypes.data_as(ctypes.c_void_p) if init is not None else None,
                                    st.ctypes.data_as(ctypes.c_void_p))
        out[name] = {"closest_nodes": round(st[0] / len(p), 2), "closest_tris": round(st[1] / len(p), 2)}
    print(json.dumps(out))


def main(n=200_000):
    if os.environ.get("PVB_STATS_SEEDED"):
        v, f = workloads.bumpy_sphere(100, 51)
        seeded("mesh10k (uniform in AABB+0.05)", v, f,
               workloads.uniform_points(n, v.min(0) - 0.05, v.max(0) + 0.05, seed=2))
        v, f = workloads.bumpy_sphere(250, 101)
        seeded("mesh50k (uniform in AABB+0.05)", v, f,
               workloads.uniform_points(n, v.min(0) - 0.05, v.max(0) + 0.05, seed=2))
        return
    v, f = workloads.bumpy_sphere(100, 51)
    run("mesh10k (uniform in AABB+0.05)", v, f, workloads.uniform_points(n, v.min(0) - 0.05, v.max(0) + 0.05, seed=2),
        nat.PVB_MESH_DEFAULT)
    v, f = workloads.bumpy_sphere(250, 101)
    run("mesh50k (uniform in AABB+0.05)", v, f, workloads.uniform_points(n, v.min(0) - 0.05, v.max(0) + 0.05, seed=2),
        nat.PVB_MESH_DEFAULT)
    g = torch.Generator().manual_seed(5)
    idx = torch.randint(0, len(v), (n,), generator=g)
    near = torch.from_numpy(np.asarray(v, dtype=np.float32))[idx] + 0.002 * torch.randn(n, 3, generator=g)
    run("C5 chamfer cloud on mesh50k (surface + 2 mm noise, unsigned)", v, f, near, 0)
    v, f = workloads.fixture_mesh("drill")
    lo, hi = v.min(0) - 0.01, v.max(0) + 0.01
    run("C1 drill (uniform in AABB+0.01)", v, f, workloads.uniform_points(n, lo, hi, seed=0), nat.PVB_MESH_DEFAULT)
    v, f = workloads.fixture_mesh("wrench")
    run("open mesh: wrench (diagonal-ray parity)", v, f,
        workloads.uniform_points(n, v.min(0) - 0.02, v.max(0) + 0.02, seed=1), nat.PVB_MESH_DEFAULT)


if __name__ == "__main__":
    main(*(int(a) for a in sys.argv[1:]))
