"""Sharding of perturbed copies (occlusion windows / FBP bands / stem masks) across the GPUs of one box.

One process per GPU (torchrun); every copy is independent given the track, so each rank sweeps a contiguous,
balanced slice (``grid.shard_range``) and a single all-gather of the per-rank probability vectors rebuilds the
full list on every rank (a few KB over NVLink).  No other communication exists on the path.
"""
from __future__ import annotations

from typing import Callable

import numpy as np

from .grid import shard_range


def world():
    """(rank, world_size) of the default process group, or (0, 1) when torch.distributed is not initialised."""
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return 0, 1


def gather_shards(local: np.ndarray, n_total: int) -> np.ndarray:
    """All-gather the ranks' contiguous float32 slices back into the full ``[n_total]`` vector (rank order)."""
    rank, ws = world()
    if ws == 1:
        return np.asarray(local, dtype=np.float32)
    import torch
    import torch.distributed as dist

    max_len = (n_total + ws - 1) // ws
    on_gpu = dist.get_backend() == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu")
    buf = torch.zeros(max_len, dtype=torch.float32, device=dev)
    buf[: len(local)] = torch.from_numpy(np.asarray(local, dtype=np.float32)).to(dev)
    out = [torch.empty_like(buf) for _ in range(ws)]
    dist.all_gather(out, buf)
    parts = []
    for r in range(ws):
        a, b = shard_range(n_total, r, ws)
        parts.append(out[r][: b - a].cpu().numpy())
    return np.concatenate(parts) if parts else np.zeros(0, np.float32)


def sharded_sweep(sweep: Callable[[np.ndarray], np.ndarray], items: np.ndarray) -> np.ndarray:
    """Run ``sweep`` on this rank's slice of ``items`` (first axis) and return the gathered full result."""
    rank, ws = world()
    n = len(items)
    a, b = shard_range(n, rank, ws)
    local = sweep(items[a:b]) if b > a else np.zeros(0, np.float32)
    return gather_shards(local, n)
