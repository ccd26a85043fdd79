"""Batch sharding across ranks: one process per GPU, contiguous slices, no data-path collective.

Every element of a batch is independent, so rank r of W owns rows
[r*n/W, (r+1)*n/W) of every input array and writes the same rows of the output.  The only collective
is the optional result gather (`gather_rows`, an all_gather of the per-rank output slices over
NCCL / NVLink when the caller wants every rank to hold the full result) and the max-over-ranks of
the elapsed time used by bench.py.  Works with any torch.distributed backend (tests use gloo).
"""
import numpy as np


def slice_bounds(n, rank, world):
    """Rows [lo, hi) owned by `rank`: same partition as run_sharded() in csrc/eccbatch.cu."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    return n * rank // world, n * (rank + 1) // world


def shard_rows(arrays, rank, world):
    n = arrays[0].shape[0]
    for a in arrays:
        if a.shape[0] != n:
            raise ValueError("count mismatch")
    lo, hi = slice_bounds(n, rank, world)
    return [a[lo:hi] for a in arrays]


def gather_rows(local, n, dist, device=None):
    """all_gather the per-rank row slices of an (n, w) uint8 result into the full array on every rank."""
    import torch

    world, rank = dist.get_world_size(), dist.get_rank()
    w = local.shape[1] if local.ndim == 2 else 1
    sizes = [slice_bounds(n, r, world) for r in range(world)]
    cap = max(hi - lo for lo, hi in sizes)
    buf = torch.zeros((cap, w), dtype=torch.uint8, device=device)
    lo, hi = sizes[rank]
    t = torch.as_tensor(np.ascontiguousarray(local).reshape(hi - lo, w))
    buf[: hi - lo] = t.to(buf.device)
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    out = torch.cat([parts[r][: sizes[r][1] - sizes[r][0]] for r in range(world)], dim=0)
    return out.cpu().numpy().reshape((n, w) if local.ndim == 2 else (n,))


def max_over_ranks(value, dist, device=None):
    import torch

    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
