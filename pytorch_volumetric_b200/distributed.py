"""Multi-GPU sharding of the query path (one process per GPU, torch.distributed).

The path is embarrassingly parallel: meshes / BVHs / tables are replicated on every rank and the work is split
into contiguous slabs, so there is no data-path collective inside a query.  The only exchange is the optional
re-assembly of the per-rank result slabs (one all-gather over NCCL / NVLink), used by the RobotSDF and chamfer
sweeps when the caller wants the full result on every rank.

  shard_range           contiguous slab [begin, end) of `n` items for (rank, world)
  sharded_query         ObjectFrameSDF over a shard of the flattened point axis (+ optional all-gather)
  sharded_robot_query   RobotSDF over a shard of the configuration batch (+ optional all-gather)
  sharded_chamfer       chamfer partial means over a shard of the cloud, all-reduced (B floats)
"""
import math

import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous split of n items; the first n % world ranks take one extra."""
    base, extra = divmod(n, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def all_gather_slabs(local, counts, group=None):
    """Concatenate per-rank slabs (dim 0, sizes `counts`) on every rank with one all_gather."""
    rank, world = _world(group)
    if world == 1:
        return local
    if len(set(counts)) == 1:
        out = torch.empty((world * counts[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    # ragged shards: pad to the largest slab, gather, strip
    m = max(counts)
    pad = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = torch.empty((world * m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return torch.cat([out[r * m:r * m + counts[r]] for r in range(world)], dim=0)


def sharded_query(sdf, points, gather=True, group=None):
    """Evaluate `sdf` on this rank's slab of the flattened points [..., N, 3].

    Returns (val, grad) for the full batch when gather=True (original leading shape), else the local slab
    (flat) together with its (begin, end)."""
    rank, world = _world(group)
    lead = tuple(points.shape[:-1])
    flat = points.reshape(-1, 3)
    n = flat.shape[0]
    begin, end = shard_range(n, rank, world)
    val, grad = sdf(flat[begin:end])
    if not gather:
        return val, grad, (begin, end)
    counts = [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]
    val = all_gather_slabs(val, counts, group)
    grad = all_gather_slabs(grad, counts, group)
    return val.reshape(lead), grad.reshape(*lead, 3)


def sharded_robot_query(robot_sdf, points, gather=True, group=None):
    """RobotSDF over this rank's contiguous slab of the (flattened) configuration batch.

    The output slab (cfg_count, P) is a contiguous block of the (|A|, P) result, so re-assembly is a plain
    concatenation on dim 0.  Returns ([A,] *B, N) / (..., 3) when gather=True, else the local slab and its range."""
    rank, world = _world(group)
    comp = robot_sdf.sdf
    n_cfg = 1 if comp.tsf_batch is None else math.prod(list(comp.tsf_batch))
    begin, end = shard_range(n_cfg, rank, world)
    P = points.reshape(-1, 3).shape[0]
    val, grad = comp.query(points, cfg_begin=begin, cfg_count=end - begin)
    val = val.reshape(end - begin, P)
    grad = grad.reshape(end - begin, P, 3)
    if not gather:
        return val, grad, (begin, end)
    counts = [shard_range(n_cfg, r, world)[1] - shard_range(n_cfg, r, world)[0] for r in range(world)]
    val = all_gather_slabs(val, counts, group)
    grad = all_gather_slabs(grad, counts, group)
    lead = tuple(points.shape[:-1])
    batch = tuple(comp.tsf_batch) if comp.tsf_batch is not None else ()
    return val.reshape(*batch, *lead), grad.reshape(*batch, *lead, 3)


def sharded_chamfer(world_to_object, points, obj_factory=None, obj_sdf=None, scale=1000., group=None):
    """batch_chamfer_dist with the cloud split over ranks; the B partial sums are all-reduced."""
    from .chamfer import batch_chamfer_dist
    rank, world = _world(group)
    n = points.shape[0]
    begin, end = shard_range(n, rank, world)
    local = batch_chamfer_dist(world_to_object, points[begin:end], obj_factory=obj_factory, obj_sdf=obj_sdf,
                               scale=scale) if end > begin else torch.zeros(world_to_object.shape[0],
                                                                            dtype=world_to_object.dtype,
                                                                            device=world_to_object.device)
    total = local * float(end - begin)
    if world > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    return total / float(n)
