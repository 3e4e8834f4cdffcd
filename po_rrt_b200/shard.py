"""Multi-GPU plumbing: independent units (edges, states, queries) are sharded contiguously over ranks,
the map / vertex set is replicated.  No collective sits on the data path; `all_gather_shards` is only for the case where a
later device-resident stage needs every rank's validity masks / neighbour lists (NCCL over NVLink on GPU tensors, gloo on
CPU tensors in the tests)."""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """contiguous [lo, hi) of n units for `rank`; sizes differ by at most one, lower ranks get the extra unit"""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_gather_shards(local, n_total, group=None):
    """concatenate the per-rank result slices (first dimension) in rank order -> tensor of n_total rows on every rank"""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]
    assert local.shape[0] == sizes[rank]
    pad = max(sizes)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], 0)


def exchange_unique_id(make_id, group=None):
    """rank 0 calls make_id() -> 128 uint8 (ncclUniqueId); every rank returns rank 0's bytes.  Any backend (gloo or nccl)."""
    rank = dist.get_rank(group)
    box = [bytes(bytearray(make_id())) if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    return box[0]


def init_comm(ctx, group=None):
    """Give the Context (one per process / GPU) an NCCL communicator spanning the torch.distributed group: porrt_prm_build and
    porrt_sssp_worlds then run sharded with their exchange step inside the library (csrc/comm.cu)."""
    import numpy as np
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return rank, world
    uid = exchange_unique_id(lambda: ctx.comm_unique_id().tolist(), group)
    ctx.comm_init(np.frombuffer(uid, dtype=np.uint8), rank, world)
    return rank, world


def max_over_ranks(value, device="cpu", group=None):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
