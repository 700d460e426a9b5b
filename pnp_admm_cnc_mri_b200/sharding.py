"""Batch sharding over the GPUs of one box.  Every image's reconstruction is independent
(the reference loop carries no state between images, S1:78), so ranks take contiguous batch shards, run
with no data-path collective, and only the final reconstructions are gathered."""
from __future__ import annotations

from typing import Callable, Tuple

import torch
import torch.distributed as dist


def shard_bounds(B: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of rank; sizes differ by at most one, earlier ranks take the extras."""
    if not (0 <= rank < world):
        raise ValueError('rank out of range')
    base, extra = divmod(B, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def reconstruct_sharded(images: torch.Tensor, solve: Callable[[torch.Tensor, int, int], torch.Tensor], group=None,
                        gather: bool = True) -> torch.Tensor:
    """images: (B,N,N) on every rank (or at least the rank's shard valid); solve(shard, lo, hi) -> (hi-lo,N,N).
    Returns the full (B,N,N) result on every rank when gather=True (one all_gather of padded shards),
    else just this rank's shard."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = images.shape[0]
    lo, hi = shard_bounds(B, world, rank)
    part = solve(images[lo:hi], lo, hi) if hi > lo else images.new_zeros((0,) + tuple(images.shape[1:]))
    if world == 1 or not gather:
        return part
    cap = (B + world - 1) // world
    buf = part.new_zeros((cap,) + tuple(images.shape[1:]))
    buf[:hi - lo] = part
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf, group=group)
    pieces = []
    for r in range(world):
        a, b = shard_bounds(B, world, r)
        pieces.append(outs[r][:b - a])
    return torch.cat(pieces, 0)
