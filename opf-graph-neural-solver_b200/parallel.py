"""Multi-GPU plumbing: the batch of grids shards across ranks (every grid is independent,
ref GNS/main.py:279-283); training adds exactly one all-reduce of the flat gradient, because
the reference's batch loss is mean(losses) (ref GNS/main.py:284)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous [start, stop) of `n_items` owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def flat_gradient(params):
    """The single buffer behind all parameter gradients when they are views of one flat
    tensor (what GNS.backward produces), else None."""
    params = [p for p in params if p.grad is not None]
    if not params:
        return None
    base = params[0].grad._base
    if base is None:
        return None
    n = 0
    for p in params:
        if p.grad._base is not base:
            return None
        n += p.grad.numel()
    return base if n == base.numel() else None


def allreduce_gradients(params, group=None, average: bool = False):
    """Sum (or average) parameter gradients over the process group with ONE collective."""
    params = [p for p in params if p.grad is not None]
    if not params or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    flat = flat_gradient(params)
    if flat is not None:
        dist.all_reduce(flat, group=group)
        if average:
            flat /= dist.get_world_size(group)
        return
    buf = torch.cat([p.grad.reshape(-1) for p in params])
    dist.all_reduce(buf, group=group)
    if average:
        buf /= dist.get_world_size(group)
    off = 0
    for p in params:
        n = p.grad.numel()
        p.grad.copy_(buf[off:off + n].view_as(p.grad))
        off += n
