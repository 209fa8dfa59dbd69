"""torchrun --nproc-per-node N scripts/peer_check.py [n_cfg n_pts steps]

Checks and times the peer-store re-assembly of a configuration-sharded RobotSDF result against (a) the unsharded
query on one GPU (bit-exact) and (b) the NCCL all-gather of the per-rank slabs.  Prints one JSON line + PEER_OK."""
import json
import os
import sys
import tempfile

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import workloads  # noqa: E402


def timed(fn, steps, warmup=5):
    for _ in range(warmup):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def main():
    n_cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    n_pts = int(sys.argv[2]) if len(sys.argv) > 2 else 20_000
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    import pytorch_volumetric_b200 as pv
    from pytorch_volumetric_b200 import distributed as pd
    d = os.path.join(tempfile.gettempdir(), f"pvb_peer_arm_{rank}")
    urdf, end = workloads.write_arm(d)
    chain = pv.build_serial_chain_from_urdf(open(urdf).read(), end).to(device="cuda")
    robot = pv.RobotSDF(chain, path_prefix=d,
                        link_sdf_cls=pv.cache_link_sdf_factory(resolution=0.02, padding=1.0, device="cuda",
                                                               cache_path=os.path.join(d, "cache.pkl")))
    robot.set_joint_configuration(workloads.arm_configurations(n_cfg).cuda())
    lo = [r[0] for r in workloads.ARM_QUERY_RANGE]; hi = [r[1] for r in workloads.ARM_QUERY_RANGE]
    pts = workloads.uniform_points(n_pts, lo, hi, seed=4).cuda()

    v_full, g_full = robot(pts)                                  # unsharded, this GPU
    res = pd.PeerResult(n_cfg, n_pts)
    v, g = pd.sharded_robot_query(robot, pts, gather="peer", result=res)
    torch.cuda.synchronize()
    ok = torch.equal(v, v_full) and torch.equal(g, g_full)
    vg, gg = pd.sharded_robot_query(robot, pts, gather=True)     # NCCL all-gather of the slabs
    ok = ok and torch.equal(vg, v_full) and torch.equal(gg, g_full)
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)

    begin, endc = pd.shard_range(n_cfg, rank, world)
    t_local = timed(lambda: robot.sdf.query(pts, cfg_begin=begin, cfg_count=endc - begin), steps)
    t_nccl = timed(lambda: pd.sharded_robot_query(robot, pts, gather=True), steps)
    t_peer = timed(lambda: pd.sharded_robot_query(robot, pts, gather="peer", result=res), steps)
    res.close()
    if rank == 0:
        remote = 16.0 * (endc - begin) * n_pts * (world - 1)
        print(json.dumps({"world": world, "n_cfg": n_cfg, "n_pts": n_pts, "bit_exact_all_ranks": bool(flag.item()),
                          "ms_no_reassembly": t_local, "ms_nccl_all_gather": t_nccl, "ms_peer_stores": t_peer,
                          "remote_bytes_per_rank": remote,
                          "nvlink_out_GBps_per_rank": remote / (t_peer * 1e-3) / 1e9}))
        if flag.item() == 1.0:
            print("PEER_OK")
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
