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
    ok = True
    results = {}
    begin, endc = pd.shard_range(n_cfg, rank, world)

    def check(tag, v, g):
        torch.cuda.synchronize()
        good = torch.equal(v, v_full) and torch.equal(g, g_full)
        results[tag + "_bit_exact"] = bool(good)
        return good

    vg, gg = pd.sharded_robot_query(robot, pts, gather=True)     # NCCL all-gather of the slabs
    ok = check("nccl", vg, gg) and ok
    peers = {}
    for backend in ("ipc", "symm"):
        try:
            peers[backend] = pd.PeerResult(n_cfg, n_pts, backend=backend)
        except Exception as e:      # noqa: BLE001
            results[backend + "_unavailable"] = repr(e)[:200]
            # every rank must take the same branch: PeerResult construction is collective and raises on all ranks
            continue
        res = peers[backend]
        for rep in range(3):        # three queries: both slots and the reuse of the first
            v, g = pd.sharded_robot_query(robot, pts, gather="peer", result=res)
            ok = check(f"peer_{backend}_{rep}", v, g) and ok
        if res.multicast:
            for rep in range(3):
                v, g = pd.sharded_robot_query(robot, pts, gather="multicast", result=res)
                ok = check(f"multicast_{rep}", v, g) and ok
        for pieces in (1, 3):
            res.dma_chunk_cfgs = 0 if pieces == 1 else max(1, -(-(endc - begin) // pieces))
            for rep in range(2):
                v, g = pd.sharded_robot_query(robot, pts, gather="dma", result=res)
                ok = check(f"dma_{backend}_{pieces}_{rep}", v, g) and ok
        results[backend + "_multicast"] = bool(res.multicast)
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)

    times = {"ms_no_reassembly": timed(lambda: robot.sdf.query(pts, cfg_begin=begin, cfg_count=endc - begin), steps),
             "ms_nccl_all_gather": timed(lambda: pd.sharded_robot_query(robot, pts, gather=True), steps)}
    for backend, res in peers.items():
        times[f"ms_peer_stores_{backend}"] = timed(
            lambda: pd.sharded_robot_query(robot, pts, gather="peer", result=res), steps)
        if res.multicast:
            times["ms_multicast_stores"] = timed(
                lambda: pd.sharded_robot_query(robot, pts, gather="multicast", result=res), steps)
        for pieces in (1, 2, 4):
            res.dma_chunk_cfgs = 0 if pieces == 1 else max(1, -(-(endc - begin) // pieces))
            times[f"ms_dma_push_{backend}_{pieces}"] = timed(
                lambda: pd.sharded_robot_query(robot, pts, gather="dma", result=res), steps)
    for res in peers.values():
        res.close()
    if rank == 0:
        remote = 16.0 * (endc - begin) * n_pts * (world - 1)
        line = {"world": world, "n_cfg": n_cfg, "n_pts": n_pts, "bit_exact_all_ranks": bool(flag.item()),
                "remote_bytes_per_rank": remote}
        line.update(times)
        line.update(results)
        for k, t in times.items():
            if k.startswith(("ms_peer", "ms_multicast", "ms_nccl", "ms_dma")):
                line[k.replace("ms_", "nvlink_ingest_GBps_")] = remote / (t * 1e-3) / 1e9
        print(json.dumps(line))
        if flag.item() == 1.0:
            print("PEER_OK")
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
