"""Named, seed-reproducible fixture states for parity claims (TEST INFRASTRUCTURE).

The reference ships no checkpoint, and its literal random init is useless for a waveform-SNR
claim: fresh `eval()` saturates 99 % of samples at +-1 because the
spectral-norm u,v are unconverged, and after warm-up the output is near-DC.  Every parity
number therefore names one of these states:

  fresh   torch.manual_seed(1234) -> Generator(h).eval()
  warm    fresh + 20 train()-mode forwards (converges sigma, fills BN running stats) -> eval()
  stress  every weight-normed conv re-drawn variance-preserving (weight_v ~ N(0, 1/fan_in),
          weight_g = ||v||, bias ~ N(0, 0.01), seed 7) + 30 train()-mode forwards -> eval();
          a real AC signal, the state to quote SNR on.

All functions work on ANY module with the reference's parameter names, i.e. both the reference
Generator (golden generation, build container) and the drop-in (tests, GPU box).  The big
tensors are regenerated from seeds; the warm-up products (BN running stats, spectral-norm u/v:
~2.6 k floats) are committed with the goldens so that the GPU box never has to run warm-ups.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch

INIT_SEED = 1234      # vec2wav/hparams.py:37
STRESS_SEED = 7
WARMUP_SEED = 99

#: name -> (hparam overrides, state, batch, frames, input seed)
GOLDEN_CASES = {
    "rb2_768_stress": (dict(resblock=1, num_wv_feat=768), "stress", 2, 24, 11),
    "rb1_768_stress": (dict(resblock='1', num_wv_feat=768), "stress", 2, 24, 12),
    "rb2_1024_warm": (dict(resblock=1, num_wv_feat=1024), "warm", 1, 17, 13),
    "rb2_768_fresh": (dict(resblock=1, num_wv_feat=768), "fresh", 1, 16, 14),
    "rb1_1024_stress": (dict(resblock='1', num_wv_feat=1024), "stress", 1, 20, 15),
    "rb2_768_stress_tiny": (dict(resblock=1, num_wv_feat=768), "stress", 3, 3, 16),
    "rb2_1024_stress": (dict(resblock=1, num_wv_feat=1024), "stress", 2, 24, 17),
}


def construct(factory, h):
    """Seeded construction: identical RNG consumption => identical init in both code bases."""
    torch.manual_seed(INIT_SEED)
    return factory(h)


def redraw_variance_preserving(gen, seed: int = STRESS_SEED) -> None:
    g = torch.Generator().manual_seed(seed)
    for _, mod in gen.named_modules():
        if not hasattr(mod, "weight_v") or not hasattr(mod, "weight_g"):
            continue
        if isinstance(mod, torch.nn.Linear):   # spectral-norm Linear also has weight_v
            continue
        v = mod.weight_v.data
        k = v.shape[2]
        if isinstance(mod, torch.nn.ConvTranspose1d):
            fan_in = v.shape[0] * k / mod.stride[0]
        else:
            fan_in = v.shape[1] * k
        v.copy_(torch.randn(v.shape, generator=g) * math.sqrt(1.0 / fan_in))
        mod.weight_g.data.copy_(v.flatten(1).norm(dim=1).view(-1, 1, 1))
        mod.bias.data.copy_(torch.randn(mod.bias.shape, generator=g) * 0.01)


def warmup(gen, steps: int, seed: int = WARMUP_SEED) -> None:
    """train()-mode forwards under no_grad: batch-stat BN updates the running stats and the
    spectral-norm power iteration converges (vec2wav/modules.py:14,16)."""
    h = gen.h
    g = torch.Generator().manual_seed(seed)
    gen.train()
    with torch.no_grad():
        for _ in range(steps):
            x = torch.randn(4, h.num_wv_feat, 50, generator=g)
            s = torch.randn(4, h.spk_dim, generator=g)
            n = torch.randn(4, h.noise_dim, generator=g)
            gen(x, s, n)
    gen.eval()


def apply_state(gen, state: str) -> None:
    if state == "fresh":
        gen.eval()
    elif state == "warm":
        warmup(gen, 20)
    elif state == "stress":
        redraw_variance_preserving(gen)
        warmup(gen, 30)
    else:
        raise ValueError(state)


def is_small_key(key: str) -> bool:
    return (".batch_nrom." in key) or key.endswith("layer.weight_u") or key.endswith("layer.weight_v")


def small_state(gen) -> Dict[str, np.ndarray]:
    return {k: v.detach().cpu().numpy().copy() for k, v in gen.state_dict().items()
            if is_small_key(k)}


def apply_state_from_golden(gen, state: str, small: Dict[str, np.ndarray]) -> None:
    """Reach the fixture state WITHOUT running warm-ups: seeded re-draw (if any) plus the
    committed warm-up products."""
    if state == "stress":
        redraw_variance_preserving(gen)
    sd = gen.state_dict()
    for k, arr in small.items():
        sd[k].copy_(torch.from_numpy(np.asarray(arr)))
    gen.eval()


def make_inputs(h, batch: int, frames: int, seed: int):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, h.num_wv_feat, frames, generator=g)
    s = torch.randn(batch, h.spk_dim, generator=g)
    n = torch.randn(batch, h.noise_dim, generator=g)
    return x, s, n


def state_checksums(gen) -> Dict[str, float]:
    """float64 sum and sum of squares per tensor: pins bit-level init parity cheaply."""
    out = {}
    for k, v in gen.state_dict().items():
        d = v.detach().double()
        out[k] = (float(d.sum()), float((d * d).sum()))
    return out


def snr_db(ref: np.ndarray, got: np.ndarray) -> float:
    """10*log10(sum ref^2 / sum (ref-got)^2) over the whole waveform."""
    ref = np.asarray(ref, dtype=np.float64)
    got = np.asarray(got, dtype=np.float64)
    err = float(((ref - got) ** 2).sum())
    sig = float((ref ** 2).sum())
    if err == 0.0:
        return float("inf")
    return 10.0 * math.log10(sig / err)
