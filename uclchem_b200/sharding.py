"""Sharding of a grid over ranks (one process per GPU).

Cells are independent, so a grid is cut into contiguous shards and each
rank integrates its own; nothing crosses NVLink during integration.  The only collective
is the final gather of results to rank 0 (``torch.distributed``; NCCL on the GPU box, gloo
in the CPU tests)."""
from __future__ import annotations


def shard_range(ncell: int, rank: int, world: int):
    """Contiguous [lo, hi) of cells owned by `rank`; sizes differ by at most one."""
    return ncell * rank // world, ncell * (rank + 1) // world


def gather_results(local, ncell: int, rank: int, world: int):
    """Gather per-rank [n_r, k] result blocks into [ncell, k] on rank 0 (None elsewhere)."""
    import torch
    import torch.distributed as dist

    if world == 1:
        return local
    sizes = [shard_range(ncell, r, world) for r in range(world)]
    nmax = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((nmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, bufs, dst=0)
    if rank != 0:
        return None
    return torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, sizes)], dim=0)
