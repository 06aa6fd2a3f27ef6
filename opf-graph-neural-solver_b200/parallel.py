"""Multi-GPU plumbing: the batch of grids shards across ranks (every grid is independent,
ref GNS/main.py:279-283); training adds exactly one all-reduce of the flat gradient, because
the reference's batch loss is mean(losses) (ref GNS/main.py:284)."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def bind_to_gpu_numa_node(device_index: int) -> bool:
    """Pin this process to the CPUs NVML reports as local to the GPU (one process per GPU): pinned host
    buffers are then allocated on the GPU's NUMA node and host<->device copies do not cross sockets.
    Returns False (and changes nothing) when NVML or the affinity call is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:
            handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = [w * 64 + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return False
        os.sched_setaffinity(0, cpus)
        return True
    except Exception:
        return False


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous [start, stop) of `n_items` owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def flat_gradient(params):
    """One tensor aliasing all parameter gradients when they sit back to back in one storage (what
    GNS.backward produces), else None."""
    params = [p for p in params if p.grad is not None]
    if not params:
        return None
    g0 = params[0].grad
    if g0.dtype != torch.float32 or not g0.is_contiguous():
        return None
    ptr, n = g0.data_ptr(), 0
    for p in params:
        g = p.grad
        if g.data_ptr() != ptr + 4 * n or not g.is_contiguous() or g.dtype != torch.float32:
            return None
        n += g.numel()
    storage = g0.untyped_storage()
    if (g0.storage_offset() + n) * 4 > storage.nbytes():
        return None
    return torch.empty(0, dtype=torch.float32, device=g0.device).set_(storage, g0.storage_offset(), (n,))


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
