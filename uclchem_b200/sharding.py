"""Sharding of a grid over ranks (one process per GPU).

Cells are independent, so every rank integrates its own cells and nothing crosses
NVLink during integration.  The only collective is the gather of results to rank 0
(``torch.distributed``; NCCL on the GPU box, gloo in the CPU tests).  Inside one process the C ABI
itself deals a grid over the devices it was bound to (``uclgpu_run_grid``, round-robin over the
cost-sorted cells); these helpers are for the one-process-per-GPU launch of ``bench.py``."""
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


def gather_rows(local, rank: int, world: int, out=None):
    """Gather equally sized per-rank blocks [n, k] to rank 0.  Returns the list of `world` blocks on
    rank 0 (in rank order; `out` may supply preallocated receive buffers) and None elsewhere."""
    import torch
    import torch.distributed as dist

    if world == 1:
        return [local]
    bufs = None
    if rank == 0:
        bufs = out if out is not None else [torch.empty_like(local) for _ in range(world)]
    dist.gather(local, bufs, dst=0)
    return bufs


def max_over_ranks(values, world: int, device=None):
    """Element-wise maximum of a small list of floats over all ranks (device-side timing is reported as
    the slowest rank's)."""
    import torch
    import torch.distributed as dist

    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]
