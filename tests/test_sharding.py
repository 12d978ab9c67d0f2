"""Host-side multi-rank logic on CPU: world_size 2 over gloo (the GPU path uses NCCL)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from wavthruvec_pytorch_b200.sharding import generate_sharded, shard_bounds, shard_by_frames, shard_sizes


def test_shard_bounds_cover_and_order():
    for total in (0, 1, 5, 16, 512):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b and c <= d
            sizes = shard_sizes(total, world)
            assert sum(sizes) == total and max(sizes) - min(sizes) <= 1
    assert shard_bounds(512, 8, 3) == (192, 256)     # cfg5: 64 utterances per GPU at N=8
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _fake_forward(x, spk, noise):
    # deterministic, per-utterance, shape-faithful stand-in for Generator.forward on CPU
    B, _, T = x.shape
    base = x.mean(dim=1, keepdim=True).repeat_interleave(320, dim=2)
    return torch.tanh(base + spk.mean(dim=1).view(B, 1, 1) - noise.mean(dim=1).view(B, 1, 1))


def _worker(rank, world, port, total, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(total, 32, 6, generator=g)
    s = torch.randn(total, 8, generator=g)
    n = torch.randn(total, 8, generator=g)
    y = generate_sharded(_fake_forward, x, s, n)
    if rank == 0:
        np.save(out_path, y.numpy())
    else:
        assert y is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [5, 4, 1])
def test_generate_sharded_world2_gloo(tmp_path, total):
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    out = str(tmp_path / "y.npy")
    mp.spawn(_worker, args=(2, port, total, out), nprocs=2, join=True)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(total, 32, 6, generator=g)
    s = torch.randn(total, 8, generator=g)
    n = torch.randn(total, 8, generator=g)
    want = _fake_forward(x, s, n).numpy()
    got = np.load(out)
    assert got.shape == (total, 1, 6 * 320)
    np.testing.assert_array_equal(got, want)


def test_shard_by_frames_balances_variable_length_items():
    rng = np.random.default_rng(1)
    for world in (1, 2, 3, 8):
        for n in (0, 1, 5, 64, 513):
            frames = rng.integers(50, 501, size=n).tolist()
            spans = shard_by_frames(frames, world)
            assert len(spans) == world and spans[0][0] == 0 and spans[-1][1] == n
            assert all(b == c for (_, b), (c, _) in zip(spans, spans[1:]))
            loads = [sum(frames[lo:hi]) for lo, hi in spans]
            if n >= 8 * world:      # no rank is off the ideal share by more than one longest utterance
                ideal = sum(frames) / world
                assert max(abs(v - ideal) for v in loads) <= max(frames)
    assert shard_by_frames([100, 100, 100, 100], 2) == [(0, 2), (2, 4)]
    assert shard_by_frames([400, 10, 10, 10], 2) == [(0, 1), (1, 4)]
    assert shard_by_frames([7], 3) in ([(0, 0), (0, 0), (0, 1)], [(0, 1), (1, 1), (1, 1)], [(0, 0), (0, 1), (1, 1)])
    with pytest.raises(ValueError):
        shard_by_frames([1, 2], 0)
