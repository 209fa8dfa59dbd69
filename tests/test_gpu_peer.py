"""GPU: multi-destination epilogue of the composed kernels (pvb_composed_query_multi) and the peer-mapped
re-assembly of a configuration-sharded RobotSDF result (distributed.PeerResult)."""
import os
import subprocess
import sys

import pytest
import torch

import workloads

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _arm(tmp_path, n_cfg):
    import pytorch_volumetric_b200 as pv
    urdf, end = workloads.write_arm(str(tmp_path))
    chain = pv.build_serial_chain_from_urdf(open(urdf).read(), end).to(device="cuda")
    s = pv.RobotSDF(chain, path_prefix=str(tmp_path),
                    link_sdf_cls=pv.cache_link_sdf_factory(resolution=0.02, padding=1.0, device="cuda",
                                                           cache_path=str(tmp_path / "arm.pkl")))
    s.set_joint_configuration(workloads.arm_configurations(n_cfg).cuda())
    return s


@pytest.mark.parametrize("n_cfg,n_pts", [(40, 4096), (40, 4100), (5, 1001)])     # cfg-major (full-sector rows; ragged last tile) / point-major
def test_multi_target_equals_single(tmp_path, n_cfg, n_pts):
    """Every destination of one multi-target launch holds exactly what the single-destination kernel writes, at the
    slab's place in the full buffer; elements outside the slab are untouched."""
    s = _arm(tmp_path, n_cfg)
    lo = [r[0] for r in workloads.ARM_QUERY_RANGE]; hi = [r[1] for r in workloads.ARM_QUERY_RANGE]
    pts = workloads.uniform_points(n_pts, lo, hi, seed=11).cuda()
    v_ref, g_ref = s.sdf.query(pts)
    v_ref = v_ref.view(n_cfg, n_pts); g_ref = g_ref.view(n_cfg, n_pts, 3)
    bufs = [(torch.full((n_cfg * n_pts,), -7.0, device="cuda"), torch.full((n_cfg * n_pts * 3,), -7.0, device="cuda"))
            for _ in range(3)]
    half = n_cfg // 2
    s.sdf.query_into(pts, bufs, cfg_begin=0, cfg_count=half)
    for v, g in bufs:
        assert torch.equal(v.view(n_cfg, n_pts)[:half], v_ref[:half])
        assert torch.equal(g.view(n_cfg, n_pts, 3)[:half], g_ref[:half])
        assert bool((v.view(n_cfg, n_pts)[half:] == -7.0).all()) and bool((g.view(n_cfg, n_pts, 3)[half:] == -7.0).all())
    s.sdf.query_into(pts, bufs[:2], cfg_begin=half, cfg_count=n_cfg - half)
    for v, g in bufs[:2]:
        assert torch.equal(v.view(n_cfg, n_pts), v_ref) and torch.equal(g.view(n_cfg, n_pts, 3), g_ref)
    with pytest.raises(ValueError):
        s.sdf.query_into(pts, [], cfg_begin=0, cfg_count=1)
    with pytest.raises(ValueError):
        s.sdf.query_into(pts, [(bufs[0][0][:10], bufs[0][1])], cfg_begin=0, cfg_count=1)


def test_peer_result_single_rank(tmp_path):
    """PeerResult without a process group: the IPC allocation, the torch views over it and gather="peer"."""
    from pytorch_volumetric_b200 import distributed as pd
    s = _arm(tmp_path, 6)
    pts = workloads.uniform_points(777, [-1, -0.5, -0.2], [0.5, 0.5, 0.8], seed=3).cuda()
    res = pd.PeerResult(6, 777)
    v, g = pd.sharded_robot_query(s, pts, gather="peer", result=res)
    v_ref, g_ref = s(pts)
    assert v.shape == v_ref.shape and g.shape == g_ref.shape
    assert torch.equal(v, v_ref) and torch.equal(g, g_ref)
    res.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs on the box")
def test_peer_reassembly_two_ranks():
    """Two processes, one per GPU: every rank ends up with the full result, bit-identical to the unsharded query
    (scripts/peer_check.py asserts and prints PEER_OK)."""
    env = dict(os.environ, PYTHONPATH=ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29631",
                          os.path.join(ROOT, "scripts", "peer_check.py")],
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0 and "PEER_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
