"""Sharding of a batch of independent units (zlib streams, gzip members, LZO blocks) over ranks.

The path has no exchange step (every stream starts with an empty window, de.ml:445; gz.ml:467-470 allocates a fresh
one per member), so multi-GPU is one process per GPU, each inflating a contiguous range of the unit index with no
data-path collective; only the small per-unit result arrays are gathered when one rank needs all of them.
"""
import numpy as np


def partition(weights, world):
    """Contiguous ranges [lo_r, hi_r) of len(weights) units for ranks 0..world-1, balanced by the prefix sum of
    `weights` (compressed + inflated bytes per unit).  Every unit belongs to exactly one rank; ranks may be empty."""
    w = np.asarray(weights, dtype=np.float64)
    n = len(w)
    if world <= 0:
        raise ValueError("world must be positive")
    cum = np.concatenate([[0.0], np.cumsum(w)])
    total = cum[-1]
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        k = int(np.searchsorted(cum, target, side="left"))
        k = min(max(k, cuts[-1]), n)
        cuts.append(k)
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def my_range(weights, rank, world):
    return partition(weights, world)[rank]


def gather_results(local, group=None):
    """All-gather per-unit int64 result rows (shape [n_local, k]) from every rank, in rank order.  Uses the default
    process group (NCCL on GPUs, gloo in the CPU tests); rows stay on the device they are on."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    local = local.contiguous()
    counts = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device), group=group)
    counts = [int(c.item()) for c in counts]
    m = max(counts) if counts else 0
    pad = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    bufs = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)
