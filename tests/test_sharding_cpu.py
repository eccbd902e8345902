"""Multi-GPU path on CPU: cells shard across ranks with no data-path collective; only the
result gather uses the process group (gloo here, NCCL on the box)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from uclchem_b200.sharding import gather_results, gather_rows, max_over_ranks, shard_range


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


def _bench_worker(rank, world, port, q):
    """What bench.py does per rank for N > 1, with the integration replaced by a function of the parameters."""
    import bench
    from uclchem_b200.params import PARAM_INDEX
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = bench.config2_params(rank=rank, world=world)
    sl = bench.slice_cells(p.shape[1], 1)
    stage = torch.zeros((len(sl), 3), dtype=torch.float64)
    stage[:, 0] = torch.from_numpy(p[PARAM_INDEX["zeta"], sl])
    stage[:, 1] = torch.from_numpy(p[PARAM_INDEX["initialdens"], sl])
    stage[:, 2] = float(rank)
    blocks = gather_rows(stage, rank, world)
    wall, kern = max_over_ranks([1.0 + rank, 10.0 - rank], world)
    if rank == 0:
        q.put(([b.numpy() for b in blocks], wall, kern))
    else:
        assert blocks is None
    dist.destroy_process_group()


def test_two_rank_bench_gather_gloo():
    """N = 2: the ranks own disjoint grids (every second zeta plane), rank 0 receives every rank's rows in rank
    order, and the timing figures are the slowest rank's."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_bench_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    blocks, wall, kern = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
    assert len(blocks) == 2 and blocks[0].shape == blocks[1].shape == (2500, 3)
    assert (blocks[0][:, 2] == 0).all() and (blocks[1][:, 2] == 1).all()
    z0, z1 = np.unique(blocks[0][:, 0]), np.unique(blocks[1][:, 0])
    assert len(z0) == len(z1) == 5 and not set(z0) & set(z1)          # a quarter of each rank's 20 zeta planes
    assert np.array_equal(np.unique(blocks[0][:, 1]), np.unique(blocks[1][:, 1]))
    assert (wall, kern) == (2.0, 10.0)
