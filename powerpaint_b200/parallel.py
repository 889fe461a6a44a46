"""Batch sharding across the GPUs of one box: requests are independent, so a batch is
split contiguously (both CFG halves of an image stay on one rank), weights are replicated, and the only
communication is a scatter of the inputs before the loop and a gather of the decoded images after it —
no collective inside the denoising loop. One process per GPU; `torch.distributed` (NCCL over
NVLink/NVSwitch on GPUs, gloo in the CPU tests) is the plumbing.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch


def shard_ranges(total: int, world: int) -> List[Tuple[int, int]]:
    """contiguous [start, end) per rank; the first `total % world` ranks take one extra request"""
    if world <= 0:
        raise ValueError("world must be positive")
    base, rem = divmod(total, world)
    out, s = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((s, s + n))
        s += n
    return out


def scatter_requests(tensors: Optional[Sequence[torch.Tensor]], shapes: Sequence[Tuple[int, ...]],
                     dtypes: Sequence[torch.dtype], device, src: int = 0, group=None) -> List[torch.Tensor]:
    """Rank `src` holds full-batch tensors [world * b, ...]; every rank receives its [b, ...] shard.
    `shapes`/`dtypes` describe ONE shard (known on all ranks)."""
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    out = []
    for k, (shape, dt) in enumerate(zip(shapes, dtypes)):
        dst = torch.empty(shape, dtype=dt, device=device)
        if rank == src:
            full = tensors[k].to(device)
            if full.shape[0] != world * shape[0]:
                raise ValueError(f"tensor {k}: batch {full.shape[0]} != world*shard {world * shape[0]}")
            chunks = [c.contiguous() for c in full.chunk(world, dim=0)]
        else:
            chunks = None
        dist.scatter(dst, chunks, src=src, group=group)
        out.append(dst)
    return out


def gather_images(images: torch.Tensor, dst: int = 0, group=None) -> Optional[torch.Tensor]:
    """every rank contributes [b, ...]; rank `dst` gets [world * b, ...] in rank order, others None"""
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bufs = [torch.empty_like(images) for _ in range(world)] if rank == dst else None
    dist.gather(images.contiguous(), bufs, dst=dst, group=group)
    return torch.cat(bufs, dim=0) if rank == dst else None


def scatter_batch(tensors: Optional[Sequence[torch.Tensor]], device, src: int = 0, group=None) -> List[torch.Tensor]:
    """Ragged form of `scatter_requests`: rank `src` holds full-batch tensors [total, ...] (all with the same `total`;
    the other ranks pass None and need to know nothing), rank r receives rows `shard_ranges(total, world)[r]` of each —
    shards differ by at most one row and may be EMPTY when total < world (the caller skips its `__call__` then).
    One object broadcast carries the shapes / dtypes; the payload travels as equal-sized (padded) scatter chunks."""
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if rank == src:
        if not tensors:
            raise ValueError("rank `src` must pass the full-batch tensors")
        totals = {int(t.shape[0]) for t in tensors}
        if len(totals) != 1:
            raise ValueError(f"all tensors must share the batch dimension, got {sorted(totals)}")
        meta = [[(tuple(t.shape), t.dtype) for t in tensors]]
    else:
        meta = [None]
    dist.broadcast_object_list(meta, src=src, group=group)
    out = []
    for k, (shape, dt) in enumerate(meta[0]):
        total = shape[0]
        ranges = shard_ranges(total, world)
        rows = max(e - s for s, e in ranges)  # == ceil(total / world)
        s, e = ranges[rank]
        dst = torch.empty((rows,) + tuple(shape[1:]), dtype=dt, device=device)
        chunks = None
        if rank == src:
            full = tensors[k].to(device)
            chunks = []
            for cs, ce in ranges:
                c = torch.zeros_like(dst)
                c[: ce - cs] = full[cs:ce]
                chunks.append(c)
        if rows > 0:
            dist.scatter(dst, chunks, src=src, group=group)
        out.append(dst[: e - s].contiguous())
    return out


def gather_batch(local: torch.Tensor, total: int, dst: int = 0, group=None) -> Optional[torch.Tensor]:
    """inverse of `scatter_batch`: rank r contributes its `shard_ranges(total, world)[r]` rows (possibly none; shape
    [0, ...] with the right trailing dimensions and dtype), rank `dst` gets [total, ...] in rank order, others None"""
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    ranges = shard_ranges(total, world)
    s, e = ranges[rank]
    if local.shape[0] != e - s:
        raise ValueError(f"rank {rank} holds {local.shape[0]} rows, its shard of {total} over {world} ranks is {e - s}")
    rows = max(ce - cs for cs, ce in ranges)
    if rows == 0:
        return local if rank == dst else None
    send = torch.zeros((rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    send[: e - s] = local
    bufs = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([b[: ce - cs] for b, (cs, ce) in zip(bufs, ranges)], dim=0)
