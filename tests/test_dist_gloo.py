"""The N>1 path on CPU: world_size-2 gloo process group, contiguous sharding + all-gather of the probability vector
(the only collective on the hot path)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from audio_deepfake_explainability_b200 import dist as xdist
from audio_deepfake_explainability_b200 import grid


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_items, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        items = np.arange(n_items * 4, dtype=np.int32).reshape(n_items, 4)
        seen = []

        def sweep(local):
            seen.append(len(local))
            return (local[:, 0].astype(np.float32) * 0.5 + rank * 1000.0)

        out = xdist.sharded_sweep(sweep, items)
        a, b = grid.shard_range(n_items, rank, world)
        q.put((rank, out.tolist(), seen, (a, b)))
    finally:
        dist.destroy_process_group()


def _run(world, n_items):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(results)


def test_sharded_sweep_world2():
    n = 11
    results = _run(2, n)
    expected = []
    for r in range(2):
        a, b = grid.shard_range(n, r, 2)
        expected += [i * 4 * 0.5 + r * 1000.0 for i in range(a, b)]
    for rank, out, seen, (a, b) in results:
        assert out == expected                      # every rank holds the full vector in window order
        assert seen == [b - a]                      # and swept only its own contiguous slice


def test_sharded_sweep_fewer_items_than_ranks():
    results = _run(2, 1)
    for rank, out, seen, (a, b) in results:
        assert out == [0.0]
        assert seen == ([1] if rank == 0 else [])


def test_single_process_is_identity():
    items = np.arange(12, dtype=np.int32).reshape(3, 4)
    out = xdist.sharded_sweep(lambda w: w[:, 1].astype(np.float32), items)
    assert out.tolist() == [1.0, 5.0, 9.0]
    assert xdist.world() == (0, 1)
