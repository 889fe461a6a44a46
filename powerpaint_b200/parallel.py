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
