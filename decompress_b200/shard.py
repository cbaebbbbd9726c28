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


def partition_cuts(weights, world):
    """`partition` for a weight tensor that lives on a device (or any torch tensor): the world + 1 cut indices as an int64
    tensor on the same device, no host round trip — the same cuts as `partition` (same prefix sums, same left-most search)."""
    import torch

    n = weights.shape[0]
    cum = torch.cat([torch.zeros(1, dtype=torch.float64, device=weights.device), torch.cumsum(weights.to(torch.float64), 0)])
    targets = cum[-1] * torch.arange(1, world, dtype=torch.float64, device=weights.device) / world
    k = torch.searchsorted(cum, targets, right=False).clamp_(max=n)
    k = torch.cummax(k, 0).values if world > 1 and k.numel() else k
    return torch.cat([torch.zeros(1, dtype=torch.int64, device=weights.device), k.to(torch.int64),
                      torch.full((1,), n, dtype=torch.int64, device=weights.device)])


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


# ---- a batch that lives on ONE rank (BASELINE config C4: "sharded 8 x B200") ------------------------------------------
# The units are independent, so the only exchange is moving bytes to where they are decoded and back: the index is
# broadcast, every rank gets the contiguous slab of compressed units `partition` gives it (point-to-point sends from the
# owner, NCCL over NVLink on GPUs, gloo in the CPU tests), decodes it, and returns its output slab to the owner.

def broadcast_index(index, owner=0, device=None, group=None):
    """`index`: int64 array [n, k] on the owner (None elsewhere).  Returns it on every rank (as a numpy array)."""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank(group)
    shape = torch.tensor(list(index.shape) if rank == owner else [0, 0], dtype=torch.int64, device=device)
    dist.broadcast(shape, src=owner, group=group)
    n, k = int(shape[0]), int(shape[1])
    t = torch.from_numpy(np.ascontiguousarray(index, dtype=np.int64)).to(device) if rank == owner else torch.empty((n, k), dtype=torch.int64, device=device)
    dist.broadcast(t, src=owner, group=group)
    return t.cpu().numpy()


def scatter_slab(blob, byte_ranges, owner=0, group=None, device=None):
    """The owner holds `blob` (1-D uint8 tensor); rank r receives blob[lo_r:hi_r].  Returns (local slab, works)."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = byte_ranges[rank]
    ops = []
    if rank == owner:
        local = blob[lo:hi]
        for r in range(world):
            if r != owner and byte_ranges[r][1] > byte_ranges[r][0]:
                ops.append(dist.P2POp(dist.isend, blob[byte_ranges[r][0]:byte_ranges[r][1]], r, group))
    else:
        local = torch.empty(hi - lo, dtype=torch.uint8, device=device)
        if hi > lo:
            ops.append(dist.P2POp(dist.irecv, local, owner, group))
    works = dist.batch_isend_irecv(ops) if ops else []
    return local, works


def gather_slab(local, out_full, byte_ranges, owner=0, group=None, pieces=None):
    """Rank r's `local` (1-D uint8) lands in out_full[lo_r:hi_r] on the owner.  `pieces`: optional list of (a, b) cuts of
    the local range, sent as separate messages (so that a piece can leave while the next one is still being decoded:
    the caller orders each send after the kernel that fills it).  Returns the works."""
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    ops = []
    if rank == owner:
        lo, hi = byte_ranges[owner]
        if local is not None and local.data_ptr() != out_full[lo:hi].data_ptr():
            out_full[lo:hi].copy_(local)
        for r in range(world):
            if r == owner:
                continue
            rlo, rhi = byte_ranges[r]
            cuts = pieces[r] if pieces else [(0, rhi - rlo)]
            for a, b in cuts:
                if b > a:
                    ops.append(dist.P2POp(dist.irecv, out_full[rlo + a:rlo + b], r, group))
    else:
        lo, hi = byte_ranges[rank]
        cuts = pieces[rank] if pieces else [(0, hi - lo)]
        for a, b in cuts:
            if b > a:
                ops.append(dist.P2POp(dist.isend, local[a:b], owner, group))
    return dist.batch_isend_irecv(ops) if ops else []
