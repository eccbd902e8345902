"""Multi-GPU path on CPU: cells shard across ranks with no data-path collective; only the
result gather uses the process group (gloo here, NCCL on the box)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from uclchem_b200.sharding import gather_results, shard_range


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 10000, 100001):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n, rank, world)
    local = torch.arange(lo, hi, dtype=torch.float64).repeat_interleave(3).reshape(-1, 3)
    full = gather_results(local, n, rank, world)
    t = torch.tensor([float(hi - lo)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put((full.numpy(), t.item()))
    dist.destroy_process_group()


def test_two_rank_gather_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n = 11
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    full, mx = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
    assert full.shape == (n, 3) and np.array_equal(full[:, 0], np.arange(n)) and mx == 6.0
