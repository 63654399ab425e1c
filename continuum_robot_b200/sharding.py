"""Member sharding across the GPUs of one box.

Members of an ensemble are independent (the reference's only parallelism is per-beam
``multiprocessing.Pool.map``, examples/beam_comparison_gravity.py:72-73), so the ensemble is cut
into contiguous member ranges, one per rank, and NO collective runs on the step path; the only
communication is one final gather of the results (NCCL over NVLink on the GPU box, gloo in the
CPU tests).  One process per GPU, launched with ``python -m torch.distributed.run``.
"""

from __future__ import annotations

from typing import List, Optional, Tuple


def shard_range(n_members: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of rank ``rank``; the first ``n_members % world`` ranks get one extra."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n_members, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(n_members: int, world: int) -> List[int]:
    return [shard_range(n_members, r, world)[1] - shard_range(n_members, r, world)[0] for r in range(world)]


def gather_members(local, n_members: int, dst: int = 0, group=None):
    """Final gather of per-member results ``local[b_local, ...]`` to rank ``dst`` in member order.

    Returns the full ``[n_members, ...]`` tensor on ``dst`` and ``None`` elsewhere.  Shards may be
    ragged; they are padded to the largest shard for the collective and trimmed afterwards.
    """
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = shard_sizes(n_members, world)
    if local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} members, expected {sizes[rank]}")
    mx = max(sizes)
    pad = local
    if local.shape[0] < mx:
        pad = torch.cat([local, local.new_zeros((mx - local.shape[0],) + tuple(local.shape[1:]))], dim=0)
    pad = pad.contiguous()
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)], dim=0)


def local_slice(array, rank: Optional[int] = None, world: Optional[int] = None):
    """Rows of a per-member host array that belong to this rank (env RANK / WORLD_SIZE by default)."""
    import os

    rank = int(os.environ.get("RANK", "0")) if rank is None else rank
    world = int(os.environ.get("WORLD_SIZE", "1")) if world is None else world
    lo, hi = shard_range(len(array), rank, world)
    return array[lo:hi]


def bind_to_gpu_numa_node(device_index: int) -> bool:
    """Pin the calling process to the CPU cores next to GPU ``device_index`` (NVML's ideal affinity)
    so that pinned host buffers allocated afterwards are first-touched on the GPU's own NUMA node:
    with 8 ranks copying state in and out concurrently, cross-socket PCIe traffic otherwise caps
    the host<->device rate.  Returns False (and changes nothing) if NVML is unavailable."""
    try:
        import pynvml

        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        return True
    except Exception:
        return False
