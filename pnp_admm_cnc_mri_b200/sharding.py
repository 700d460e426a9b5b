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
                        gather: bool = True, out_dtype: torch.dtype = None, out_device=None) -> torch.Tensor:
    """images: (B,N,N) on every rank (or at least the rank's shard valid); solve(shard, lo, hi) -> (hi-lo,N,N).
    Returns the full (B,N,N) result on every rank when gather=True (one all_gather of padded shards),
    else just this rank's shard.  `out_dtype` / `out_device`: what `solve` returns (default: float32 on the current
    CUDA device when the process group is NCCL, else the input's own); a rank whose shard is empty (B < world size)
    must contribute a buffer of the same dtype and device as the others or the collective fails."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = images.shape[0]
    lo, hi = shard_bounds(B, world, rank)
    if hi > lo:
        part = solve(images[lo:hi], lo, hi)
        if (out_dtype is not None and part.dtype != out_dtype) or (out_device is not None and part.device != torch.device(out_device)):
            raise ValueError(f'solve returned {part.dtype} on {part.device}, expected {out_dtype} on {out_device}')
    else:
        nccl = dist.is_initialized() and dist.get_backend(group) == 'nccl'
        dt = out_dtype if out_dtype is not None else (torch.float32 if nccl else images.dtype)
        dev = out_device if out_device is not None else (torch.device('cuda', torch.cuda.current_device()) if nccl else images.device)
        part = torch.zeros((0,) + tuple(images.shape[1:]), dtype=dt, device=dev)
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
