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


def _gather_padded(local: torch.Tensor, sizes: List[int], rank: int, dst: int, group,
                   async_op: bool = False):
    """Gather of per-rank [n_r, 1, L] waveforms padded to the largest shard (a plain gather works on every
    backend).  Returns (parts or None, work or None)."""
    pad = max(sizes)
    if local.shape[0] == pad:
        buf = local
    else:
        buf = local.new_zeros((pad,) + tuple(local.shape[1:]))
        buf[: local.shape[0]] = local
    parts = [torch.empty_like(buf) for _ in sizes] if rank == dst else None
    work = dist.gather(buf, parts, dst=dst, group=group, async_op=async_op)
    return parts, work


def generate_sharded(forward: Callable[[torch.Tensor, torch.Tensor, torch.Tensor], torch.Tensor],
                     x: torch.Tensor, spk: torch.Tensor, noise: torch.Tensor,
                     group: Optional[dist.ProcessGroup] = None, dst: int = 0,
                     samples_per_frame: int = 320, total: Optional[int] = None) -> Optional[torch.Tensor]:
    """Each rank runs `forward` on its contiguous slice of the global batch; rank `dst` returns the
    full [B, 1, 320 T] waveform in batch order, the others return None.  Works for uneven and empty
    shards.  Two calling conventions:
      * `total` is None: x / spk / noise are the GLOBAL batch (every rank holds it, or at least its own
        slice of it) and the slice is taken here;
      * `total` = the global batch size: x / spk / noise are ALREADY this rank's slice
        (`shard_bounds(total, world, rank)` rows) -- no rank ever holds more than its shard."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if total is None:
        total = x.shape[0]
        lo, hi = shard_bounds(total, world, rank)
        x, spk, noise = x[lo:hi], spk[lo:hi], noise[lo:hi]
    else:
        lo, hi = shard_bounds(total, world, rank)
        if x.shape[0] != hi - lo:
            raise ValueError(f"rank {rank} of {world} owns {hi - lo} of {total} utterances, got {x.shape[0]}")
    length = x.shape[2] * samples_per_frame
    if hi > lo:
        local = forward(x, spk, noise)
    else:
        local = x.new_zeros((0, 1, length), dtype=torch.float32)
    if world == 1:
        return local
    sizes = shard_sizes(total, world)
    parts, _ = _gather_padded(local, sizes, rank, dst, group)
    if rank == dst:
        return torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0)
    return None


class ShardedStream:
    """Steady-state sharded generation: the output gather of step i runs (asynchronously, on the
    backend's own stream) while step i + 1 is generated.  No collective touches the forward itself.

        st = ShardedStream(forward, total)
        for batch in batches:                 # each rank passes ITS slice of every global batch
            done = st.step(x_local, spk_local, noise_local)     # rank dst: waveform of the PREVIOUS step (or None)
        last = st.flush()

    Buffers are double-buffered by the caller's forward (each call returns a fresh tensor)."""

    def __init__(self, forward, total: int, group: Optional[dist.ProcessGroup] = None, dst: int = 0):
        self.forward, self.total, self.group, self.dst = forward, int(total), group, dst
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.sizes = shard_sizes(self.total, self.world)
        self._pending = None

    def _finish(self):
        if self._pending is None:
            return None
        parts, work, local = self._pending
        self._pending = None
        if self.world == 1:
            return local
        work.wait()
        if self.rank == self.dst:
            return torch.cat([p[:n] for p, n in zip(parts, self.sizes)], dim=0)
        return None

    def step(self, x, spk, noise):
        local = self.forward(x, spk, noise)
        done = self._finish()
        if self.world == 1:
            self._pending = (None, None, local)
        else:
            parts, work = _gather_padded(local, self.sizes, self.rank, self.dst, self.group, async_op=True)
            self._pending = (parts, work, local)
        return done

    def flush(self):
        return self._finish()
