"""Multi-rank host logic on CPU: world_size-2 (and 3) gloo groups exercise the contiguous-slice
sharding, the optional result gather and the max-over-ranks reduction that bench.py uses at N > 1.
The per-rank "device" work is done by the oracle here (no GPU in this container)."""
import os
import socket

import numpy as np
import pytest

from eccoxide_b200.shard import slice_bounds


def test_slice_bounds_partition_exactly():
    for n in (0, 1, 2, 7, 64, 1000, 65536 + 3):
        for world in (1, 2, 3, 4, 8):
            b = [slice_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1
    with pytest.raises(ValueError):
        slice_bounds(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from eccoxide_b200.shard import gather_rows, max_over_ranks, shard_rows
        from oracle import coracle as C

        g = np.random.Generator(np.random.Philox(1234))  # same inputs on every rank
        k = g.integers(0, 256, size=(n, 32), dtype=np.uint8)
        u = g.integers(0, 256, size=(n, 32), dtype=np.uint8)
        mine = shard_rows([k, u], rank, world)
        local = C.x25519(mine[0], mine[1])
        full = gather_rows(local, n, dist)
        t = max_over_ranks(1.0 + rank, dist)
        dist.barrier()
        if rank == 0:
            q.put((full.tobytes(), t))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 101), (3, 64)])
def test_sharded_batch_equals_single_rank(world, n, coracle):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    full, t = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = np.random.Generator(np.random.Philox(1234))
    k = g.integers(0, 256, size=(n, 32), dtype=np.uint8)
    u = g.integers(0, 256, size=(n, 32), dtype=np.uint8)
    assert full == coracle.x25519(k, u).tobytes()
    assert t == float(world)
