"""Small multi-destination composed query for compute-sanitizer (memcheck / racecheck):

    compute-sanitizer --tool memcheck python scripts/sanitize_multi.py

4 grid sub-SDFs (probe mesh, coarse table so the build stays cheap under the sanitizer) + 1 sphere, 40 and 5
configurations, sizes that take the full-sector row path, the ragged last tile and the point-major kernels."""
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads  # noqa: E402
import pytorch_volumetric_b200 as pv  # noqa: E402


def main():
    v, f = workloads.fixture_mesh("probe")
    obj = pv.MeshObjectFactory("probe", mesh=(v, f))
    cache = os.path.join(tempfile.mkdtemp(), "c.pkl")
    grid = pv.CachedSDF("probe", 0.02, obj.bounding_box(padding=0.05), pv.MeshSDF(obj), device="cuda", cache_path=cache)
    subs = [grid, grid, pv.SphereSDF(0.05), grid, grid]
    ok = True
    for n_cfg, n_pts in ((40, 4096), (40, 4100), (5, 1001)):
        tm = workloads.random_rigid(len(subs) * n_cfg, seed=7, t_range=0.2).cuda()
        comp = pv.ComposedSDF(subs, pv.Transform3d(matrix=tm))
        comp.set_transforms(pv.Transform3d(matrix=tm), batch_dim=(n_cfg,))
        pts = workloads.uniform_points(n_pts, [-0.4] * 3, [0.4] * 3, seed=1).cuda()
        v_ref, g_ref = comp.query(pts)
        bufs = [(torch.zeros(n_cfg * n_pts, device="cuda"), torch.zeros(n_cfg * n_pts * 3, device="cuda")) for _ in range(3)]
        half = n_cfg // 2
        comp.query_into(pts, bufs, cfg_begin=0, cfg_count=half)
        comp.query_into(pts, bufs, cfg_begin=half, cfg_count=n_cfg - half)
        torch.cuda.synchronize()
        for bv, bg in bufs:
            ok = ok and torch.equal(bv, v_ref) and torch.equal(bg.view(-1, 3), g_ref)
    print("SANITIZE_MULTI_OK" if ok else "SANITIZE_MULTI_MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
