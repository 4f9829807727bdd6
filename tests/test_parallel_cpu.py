"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: row sharding, batch sharding, the
max-over-ranks timing reduction and chain gathering."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from linna_b200 import parallel


def test_shard_rows_partition():
    for n in (0, 1, 7, 100000, 10001):
        for ws in (1, 2, 3, 8):
            spans = [parallel.shard_rows(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, ws, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        r, w = parallel.world()
        assert (r, w) == (rank, ws)
        # independent walkers: every rank evaluates its own slice, nothing is exchanged on the data path
        n = 1001
        lo, hi = parallel.shard_rows(n, rank, ws)
        local = torch.arange(lo, hi, dtype=torch.float32)[:, None] * torch.ones(1, 3)
        full = parallel.gather_rows(local)
        assert full.shape == (n, 3) and torch.equal(full[:, 0], torch.arange(n, dtype=torch.float32))
        # timing rule: max over ranks
        t = parallel.max_over_ranks(1.0 + rank)
        assert t == float(ws)
        # data-parallel batch: same shuffle on every rank, disjoint rows, mean gradient == full-batch gradient
        g = torch.Generator().manual_seed(7)
        idx = torch.randperm(20, generator=g)[:10]
        mine = parallel.batch_shard(idx, rank, ws)
        x = torch.arange(20, dtype=torch.float64)
        grad_local = x[mine].mean()[None].clone()
        dist.all_reduce(grad_local, op=dist.ReduceOp.SUM)
        grad_local /= ws
        assert abs(float(grad_local) - float(x[idx].mean())) < 1e-12
        q.put((rank, "ok"))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=5) for _ in range(2))
    assert got == [(0, "ok"), (1, "ok")]
