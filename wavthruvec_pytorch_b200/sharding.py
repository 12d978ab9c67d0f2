"""Batch sharding of independent utterances across ranks (one process per GPU).

In eval() every utterance's forward is independent (BN uses running statistics, CBN is
per-sample: vec2wav/modules.py:23-28), so the path shards by contiguous batch slices with NO
collective inside the forward.  `torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is used
only to gather the output waveforms -- and, in bench.py, timing scalars.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of `total` utterances owned by `rank`; the first `total % world`
    ranks hold one extra utterance, trailing ranks may be empty."""
    if world < 1 or not 0 <= rank < world or total < 0:
        raise ValueError(f"bad shard request total={total} world={world} rank={rank}")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(total: int, world: int) -> List[int]:
    return [hi - lo for lo, hi in (shard_bounds(total, world, r) for r in range(world))]


def shard_by_frames(frames, world: int) -> List[Tuple[int, int]]:
    """Contiguous item ranges [lo, hi), one per rank, balanced by total FRAMES instead of item count --
    for variable-length corpora (the file pipeline): rank r's range ends where the running frame count
    first reaches (r + 1) / world of the total.  Every rank computes the same partition from the same
    frame counts; ranges may be empty when there are fewer items than ranks."""
    if world < 1:
        raise ValueError(f"bad world size {world}")
    frames = [int(f) for f in frames]
    if any(f < 0 for f in frames):
        raise ValueError("negative frame count")
    total = sum(frames)
    bounds, lo, acc, i = [], 0, 0, 0
    for r in range(world):
        target = total * (r + 1) / world
        while i < len(frames) and (acc + frames[i] / 2.0 <= target or r == world - 1):
            acc += frames[i]
            i += 1
        bounds.append((lo, i))
        lo = i
    return bounds


def generate_sharded(forward: Callable[[torch.Tensor, torch.Tensor, torch.Tensor], torch.Tensor],
                     x: torch.Tensor, spk: torch.Tensor, noise: torch.Tensor,
                     group: Optional[dist.ProcessGroup] = None, dst: int = 0,
                     samples_per_frame: int = 320) -> Optional[torch.Tensor]:
    """Every rank holds the same global batch (or at least its own slice of it); each runs
    `forward` on its slice; rank `dst` returns the full [B, 1, 320 T] waveform in batch order,
    the others return None.  Works for uneven and empty shards."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    total = x.shape[0]
    lo, hi = shard_bounds(total, world, rank)
    length = x.shape[2] * samples_per_frame
    if hi > lo:
        local = forward(x[lo:hi], spk[lo:hi], noise[lo:hi])
    else:
        local = x.new_zeros((0, 1, length), dtype=torch.float32)
    if world == 1:
        return local
    sizes = shard_sizes(total, world)
    # pad to the largest shard so that a plain gather works on every backend
    pad = max(sizes)
    buf = local.new_zeros((pad, 1, length))
    buf[: hi - lo] = local
    if rank == dst:
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.gather(buf, parts, dst=dst, group=group)
        return torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0)
    dist.gather(buf, None, dst=dst, group=group)
    return None
