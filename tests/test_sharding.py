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


def _worker_local(rank, world, port, total, steps, out_path):
    """Local-slice convention + the pipelined gather (ShardedStream): every rank only ever holds its shard."""
    from wavthruvec_pytorch_b200.sharding import ShardedStream
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(total, world, rank)
    outs = []
    st = ShardedStream(_fake_forward, total)
    for k in range(steps):
        g = torch.Generator().manual_seed(k)
        x = torch.randn(total, 32, 6, generator=g)
        s = torch.randn(total, 8, generator=g)
        n = torch.randn(total, 8, generator=g)
        if k == 0:
            y0 = generate_sharded(_fake_forward, x[lo:hi], s[lo:hi], n[lo:hi], total=total)
        done = st.step(x[lo:hi], s[lo:hi], n[lo:hi])
        assert (done is None) == (k == 0 or rank != 0)
        if done is not None:
            outs.append(done)
    last = st.flush()
    if rank == 0:
        outs.append(last)
        np.save(out_path, np.stack([y0.numpy()] + [o.numpy() for o in outs]))
    else:
        assert last is None and y0 is None
    with pytest.raises(ValueError):
        generate_sharded(_fake_forward, torch.zeros(hi - lo + 1, 32, 6), torch.zeros(hi - lo + 1, 8),
                         torch.zeros(hi - lo + 1, 8), total=total)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [5, 2])
def test_sharded_stream_and_local_slices_world2_gloo(tmp_path, total):
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    out = str(tmp_path / "y.npy")
    steps = 3
    mp.spawn(_worker_local, args=(2, port, total, steps, out), nprocs=2, join=True)
    got = np.load(out)
    assert got.shape == (steps + 1, total, 1, 6 * 320)
    for k in range(steps):
        g = torch.Generator().manual_seed(k)
        x = torch.randn(total, 32, 6, generator=g)
        s = torch.randn(total, 8, generator=g)
        n = torch.randn(total, 8, generator=g)
        want = _fake_forward(x, s, n).numpy()
        np.testing.assert_array_equal(got[k + 1], want)
        if k == 0:
            np.testing.assert_array_equal(got[0], want)


# ----------------------------------------------------------------------------------------------
# the product multi-GPU path on real devices: 2 ranks over NCCL, the native forward per shard
# ----------------------------------------------------------------------------------------------
def _worker_nccl(rank, world, port, out_path):
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent))
    from helpers import dropin_in_state, load_golden
    from oracle import fixtures
    from wavthruvec_pytorch_b200.sharding import ShardedStream
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    g = load_golden("rb2_768_stress")
    gen = dropin_in_state(g, cache=False).to(dev).eval()
    gen.precision = "bf16"
    total, frames = 5, 37                                   # uneven shards: 3 + 2 utterances
    x, s, n = fixtures.make_inputs(gen.h, total, frames, 77)
    lo, hi = shard_bounds(total, world, rank)
    with torch.no_grad():
        fwd = lambda a, b, c: gen(a, b, c)
        y = generate_sharded(fwd, x[lo:hi].to(dev), s[lo:hi].to(dev), n[lo:hi].to(dev), total=total)
        st = ShardedStream(fwd, total)
        for _ in range(3):
            st.step(x[lo:hi].to(dev), s[lo:hi].to(dev), n[lo:hi].to(dev))
        y2 = st.flush()
        if rank == 0:
            whole = gen(x.to(dev), s.to(dev), n.to(dev))    # the single-GPU answer for the same batch
            torch.cuda.synchronize()
            np.save(out_path, np.stack([y.cpu().numpy(), y2.cpu().numpy(), whole.cpu().numpy()]))
        else:
            assert y is None and y2 is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_generate_sharded_two_gpus_nccl_matches_single_gpu(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    out = str(tmp_path / "y.npy")
    mp.spawn(_worker_nccl, args=(2, port, out), nprocs=2, join=True)
    sharded, streamed, whole = np.load(out)
    assert whole.shape == (5, 1, 37 * 320) and np.abs(whole).max() > 1e-3
    np.testing.assert_array_equal(sharded, whole)           # utterances are independent: bit-exact
    np.testing.assert_array_equal(streamed, whole)
