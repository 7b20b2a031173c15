"""Sharding rule (host logic) + its use under torch.distributed with the gloo backend, world 2."""
import os
import socket
import sys

import pytest

import cases
from inflatox_b200.sharding import shard


@pytest.mark.parametrize("n_rows,n_vec,world", [(16384, 1, 8), (1000, 1, 3), (7, 1, 8), (1024, 1024, 8),
                                                 (64, 3, 8), (5, 1, 1)])  # fmt: skip
def test_shards_tile_the_work_exactly(n_rows, n_vec, world):
    cells = set()
    for r in range(world):
        (a, b), (c, d) = shard(n_rows, n_vec, r, world)
        assert 0 <= a <= b <= n_rows and 0 <= c <= d <= n_vec
        for row in range(a, b) if n_rows < 2000 else (a, b - 1) if b > a else ():
            for v in range(c, d) if n_vec < 100 else (c, d - 1) if d > c else ():
                assert (row, v) not in cells
                cells.add((row, v))
    rows = sum(shard(n_rows, n_vec, r, world)[0][1] - shard(n_rows, n_vec, r, world)[0][0] for r in range(world))
    vecs = sum(shard(n_rows, n_vec, r, world)[1][1] - shard(n_rows, n_vec, r, world)[1][0] for r in range(world))
    if n_vec >= world and world > 1 and n_vec > 1:
        assert vecs == n_vec
    else:
        assert rows == n_rows


def _worker(rank, world, port, n_rows, q):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    (a, b), _ = shard(n_rows, 1, rank, world)
    mine = torch.zeros(n_rows, dtype=torch.int64)
    mine[a:b] = 1
    dist.all_reduce(mine)  # every row owned exactly once across ranks
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # the timing reduction bench.py does
    dist.barrier()
    q.put((rank, bool((mine == 1).all()), float(t.item())))
    dist.destroy_process_group()


def test_row_sharding_under_gloo_world_2():
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 1001, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == [(0, True, 2.0), (1, True, 2.0)]
