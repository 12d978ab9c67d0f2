"""Variable-length batched generation: host logic on CPU, the oracle pins
the halo; `-m gpu` tests compare v2w_forward_ragged with the padded forward and with the oracle."""
from __future__ import annotations

import ctypes as C
import itertools
import struct

import numpy as np
import pytest
import torch

from helpers import dropin_in_state, hparams, load_golden, t
from oracle import fixtures, vec2wav_torch_cpu
from wavthruvec_pytorch_b200 import native, varlen


# ----------------------------------------------------------------------------- host logic (CPU)
def _cost(lengths, buckets, launch_cost):
    return sum(len(idx) * frames for idx, frames in buckets) + launch_cost * len(buckets)


@pytest.mark.parametrize("seed", range(6))
def test_plan_buckets_is_a_partition_and_optimal(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 8))
    lengths = [int(v) for v in rng.integers(1, 400, n)]
    halo, k, lc = 10, 3, 64
    buckets = varlen.plan_buckets(lengths, halo, max_buckets=k, launch_cost=lc, round_to=1)
    assert sorted(i for idx, _ in buckets for i in idx) == list(range(n))
    t_max = max(lengths)
    for idx, frames in buckets:
        longest = max(lengths[i] for i in idx)
        assert frames == min(longest + halo, t_max) and frames >= longest
    # brute force over every assignment of the sorted utterances to <= k contiguous groups
    order = sorted(range(n), key=lambda i: -lengths[i])
    best = None
    for cuts in itertools.chain.from_iterable(itertools.combinations(range(1, n), c) for c in range(min(k, n))):
        edges = [0, *cuts, n]
        groups = [order[a:b] for a, b in zip(edges, edges[1:])]
        c = sum(len(g) * min(lengths[g[0]] + halo, t_max) for g in groups) + lc * len(groups)
        best = c if best is None else min(best, c)
    assert _cost(lengths, buckets, lc) == best


def test_plan_buckets_edge_cases():
    assert varlen.plan_buckets([], 10) == []
    assert varlen.plan_buckets([50], 10) == [([0], 50)]            # never past the batch's longest
    one = varlen.plan_buckets([200] * 16, 10)
    assert len(one) == 1 and sorted(one[0][0]) == list(range(16)) and one[0][1] == 200
    with pytest.raises(ValueError):
        varlen.plan_buckets([5, 0], 10)
    # a long-tailed batch: bucketing computes far fewer frames than padding to the longest
    lengths = [500] + [60] * 15
    buckets = varlen.plan_buckets(lengths, 10)
    padded, ragged = varlen.padded_frames(lengths, buckets)
    assert padded == 16 * 500 and ragged == 500 + 15 * 72


def _halo(**over):
    lib = native.load()
    cfg = native.make_config(hparams(**over))
    m = C.c_void_p()
    native.check(lib.v2w_model_create(C.byref(cfg), C.byref(m)))
    try:
        return int(lib.v2w_model_halo_frames(m))
    finally:
        lib.v2w_model_destroy(m)


def test_halo_frames_by_hand():
    # conv_post 3 | per stage +20 (ResBlock2: (1 + 3) * 5) | ups (u, p): ceil((R + p) / u) | conv_pre 3
    # L4: 23 -> ups4 (2,1): 12 -> 32 -> ups3 (2,1): 17 -> 37 -> ups2 (4,2): 10 -> 30 -> ups1 (4,2): 8
    # -> 28 -> ups0 (5,3): 7 -> +3
    assert _halo() == 10
    # ResBlock1: (1+1 + 3+1 + 5+1) * 5 = 60 per stage:
    # 63 -> 32 -> 92 -> 47 -> 107 -> 28 -> 88 -> 23 -> 83 -> 18 -> 21
    assert _halo(resblock='1') == 21


@pytest.mark.parametrize("name", ["rb2_768_stress", "rb1_768_stress"])
def test_oracle_pins_the_halo(name):
    """The reference pads features with zeros to the longest utterance of the batch and generates
    the tail.  Claim: the first T*320 samples do not change when the padded length is cut to
    T + halo -- checked on the CPU port of the reference, in fp64 so that only real dependence
    shows.  (The halo is the structural receptive field; the coupling through a chain of edge taps
    decays fast, so shorter cuts differ by little -- but cutting at T itself, i.e. generating each
    utterance un-padded, is a different function: the padded region is not silent, biases and the
    conditional affine make it non-zero from conv_pre on.)"""
    g = load_golden(name)
    gen = dropin_in_state(g)
    h = gen.h
    halo = _halo(**g.overrides)
    sd = {k: v.double() for k, v in gen.state_dict().items()}
    rng = np.random.default_rng(3)
    T, T_max = 9, 9 + 2 * halo + 5
    x = np.zeros((1, h.num_wv_feat, T_max))
    x[:, :, :T] = rng.standard_normal((1, h.num_wv_feat, T))
    spk, noise = t(g.spk[:1]).double(), t(g.noise[:1]).double()

    def run(frames):
        with torch.no_grad():
            y = vec2wav_torch_cpu.generator_forward(sd, h, t(x[:, :, :frames]), spk, noise)
        return y.numpy()[0, 0, :T * 320]

    full = run(T_max)
    assert np.abs(run(T + halo) - full).max() < 1e-12
    assert np.abs(run(T) - full).max() > 1e-6


def test_write_wav_header(tmp_path):
    pcm = torch.tensor([0, 1, -1, 32767, -32768], dtype=torch.int16)
    path = tmp_path / "a.wav"
    varlen.write_wav(path, pcm)
    raw = path.read_bytes()
    assert raw[:4] == b"RIFF" and raw[8:16] == b"WAVEfmt " and raw[36:40] == b"data"
    assert struct.unpack("<I", raw[4:8])[0] == 36 + 10 and struct.unpack("<I", raw[40:44])[0] == 10
    fmt = struct.unpack("<IHHIIHH", raw[16:36])
    assert fmt == (16, 1, 1, 16000, 32000, 2, 16)
    assert np.frombuffer(raw[44:], dtype="<i2").tolist() == pcm.tolist()


# ------------------------------------------------------------------------------------ GPU parity
LENGTHS = [137, 12, 150, 1, 29, 150, 8]


def _varlen_case(g, gen, seed=11):
    h = gen.h
    rng = np.random.default_rng(seed)
    feats = [rng.standard_normal((n, h.num_wv_feat)).astype(np.float32) for n in LENGTHS]
    B = len(LENGTHS)
    spk = np.resize(g.spk, (B, g.spk.shape[1])).astype(np.float32) * np.linspace(0.5, 1.5, B, dtype=np.float32)[:, None]
    noise = np.resize(g.noise, (B, g.noise.shape[1])).astype(np.float32)
    x_pad = np.zeros((B, h.num_wv_feat, max(LENGTHS)), np.float32)      # collate_fn_tensor
    for b, f in enumerate(feats):
        x_pad[b, :, :f.shape[0]] = f.T
    return feats, spk, noise, x_pad


@pytest.mark.gpu
@pytest.mark.parametrize("name,precision", [("rb2_768_stress", "bf16"), ("rb1_768_stress", "fp16"),
                                            ("rb2_768_stress", "fp32"), ("rb2_768_stress", "fp32tc")])
def test_ragged_equals_padded_forward(name, precision):
    """generate_varlen == padded forward truncated, sample for sample, on the same engine."""
    g = load_golden(name)
    gen = dropin_in_state(g, cache=False).to("cuda:0")
    gen.precision = precision
    feats, spk, noise, x_pad = _varlen_case(g, gen)
    with torch.no_grad():
        y_pad = gen(t(x_pad).cuda(), t(spk).cuda(), t(noise).cuda()).cpu().numpy()
    eng = gen.engine()
    assert len(varlen.plan_buckets(LENGTHS, eng.halo_frames(), 3)) > 1
    # one launch sequence with tile skipping, and three length buckets
    for buckets in (1, 3):
        outs = varlen.generate_varlen(gen, [t(f).cuda() for f in feats], t(spk).cuda(), t(noise).cuda(),
                                      max_buckets=buckets)
        assert [o.numel() for o in outs] == [n * 320 for n in LENGTHS]
        for b, o in enumerate(outs):
            want = y_pad[b, 0, :LENGTHS[b] * 320]
            assert np.abs(o.cpu().numpy() - want).max() <= 1e-6, f"{buckets} bucket(s), utterance {b}"
    # 16-bit PCM is the rounded fp32 result
    pcm = varlen.generate_varlen(gen, [t(f).cuda() for f in feats], t(spk).cuda(), t(noise).cuda(),
                                 pcm16=True, max_buckets=3)
    for o, q in zip(outs, pcm):
        want = np.clip(np.rint(o.cpu().numpy().astype(np.float32) * np.float32(32768.0)), -32768, 32767)
        assert q.dtype == torch.int16 and np.array_equal(q.cpu().numpy(), want.astype(np.int16))


@pytest.mark.gpu
def test_ragged_fp32_against_oracle():
    """fp32 plan, ragged path, against the CPU port of the reference on the zero-padded batch."""
    g = load_golden("rb2_768_stress")
    gen = dropin_in_state(g, cache=False)
    feats, spk, noise, x_pad = _varlen_case(g, gen)
    with torch.no_grad():
        ref = vec2wav_torch_cpu.generator_forward(gen.state_dict(), gen.h, t(x_pad), t(spk), t(noise)).numpy()
    gen = gen.to("cuda:0")
    gen.precision = "fp32"
    outs = varlen.generate_varlen(gen, [t(f).cuda() for f in feats], t(spk).cuda(), t(noise).cuda())
    for b, o in enumerate(outs):
        err = np.abs(o.cpu().numpy() - ref[b, 0, :LENGTHS[b] * 320]).max()
        assert err <= 1e-4, f"utterance {b}: max-abs {err}"


@pytest.mark.gpu
def test_ragged_argument_errors():
    g = load_golden("rb2_768_stress")
    gen = dropin_in_state(g, cache=False).to("cuda:0")
    eng = gen.engine()
    rows = torch.zeros(30, gen.h.num_wv_feat, device="cuda:0")
    spk, noise = t(g.spk[:2]).cuda(), t(g.noise[:2]).cuda()
    with pytest.raises(RuntimeError, match="frames"):
        eng.forward_ragged(rows, [20, 10], spk, noise, plan_frames=16)      # longer than the plan
    with pytest.raises(RuntimeError):
        eng.forward_ragged(rows, [20, 5], spk, noise, plan_frames=24)       # rows != sum(lengths)
    with pytest.raises(RuntimeError):
        eng.forward_ragged(rows.cpu(), [20, 10], spk, noise, plan_frames=24)   # no CPU fallback
