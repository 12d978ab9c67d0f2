"""Pin parity AT THE BENCHMARKED SIZES: run the UNMODIFIED reference (build container only) on the
stress fixture at the BASELINE.json configurations and store a compact pin per case.

    python -m oracle.make_config_pins [case ...]        # default: every case of CONFIG_PINS

The full waveforms are too big to commit (cfg5 per GPU: 64 x 160000 fp32 = 41 MB), so each
tests/golden/<case>.npz holds
  * `idx`, `y_s`     every PIN_STRIDE-th sample of the flattened [B, L] waveform (PIN_STRIDE is prime:
                     the samples walk through every position inside a window tile and every utterance),
  * `sum`, `sumsq`   per-utterance float64 sum and sum of squares of the fp32 waveform,
  * `full_b`, `y_full`  FULL_UTTS whole utterances (first and last of the batch),
  * the warm-up products (BN running statistics, spectral-norm u / v) of the fixture state.
Inputs are regenerated from the seed (oracle.fixtures.make_inputs).  The GPU tests
(tests/test_config_size.py) and bench.py's `parity` self-check compare against these pins.
"""
from __future__ import annotations

import json
import sys
import time
import warnings
from pathlib import Path

import numpy as np
import torch

from . import fixtures

OUT_DIR = Path(__file__).resolve().parent.parent / "tests" / "golden"
PIN_STRIDE = 997

#: case -> (hparam overrides, state, batch, frames, input seed) at the sizes of BASELINE.json configs
CONFIG_PINS = {
    "cfg1_stress": (dict(resblock=1, num_wv_feat=768), "stress", 1, 100, 20),
    "cfg2_stress": (dict(resblock=1, num_wv_feat=768), "stress", 16, 200, 21),
    "cfg3_stress": (dict(resblock=1, num_wv_feat=1024), "stress", 32, 300, 22),
    "cfg4_stress": (dict(resblock=1, num_wv_feat=768), "stress", 1, 1500, 108),   # a seed whose single utterance is an AC signal (most are near-DC)
    "cfg5gpu_stress": (dict(resblock=1, num_wv_feat=768), "stress", 64, 500, 24),
    "cfg2_rb1_stress": (dict(resblock='1', num_wv_feat=768), "stress", 16, 200, 25),
}


def pin_indices(total: int) -> np.ndarray:
    return np.arange(PIN_STRIDE // 2, total, PIN_STRIDE, dtype=np.int64)


def full_utterances(batch: int):
    return sorted({0, batch - 1})


def summarise(y: np.ndarray) -> dict:
    """The pin of a [B, 1, L] fp32 waveform (also used by the checkers on the GPU side)."""
    b = y.shape[0]
    flat = np.ascontiguousarray(y, dtype=np.float32).reshape(b, -1)
    idx = pin_indices(flat.size)
    d = flat.astype(np.float64)
    fb = full_utterances(b)
    return {"idx": idx, "y_s": flat.reshape(-1)[idx].copy(), "sum": d.sum(axis=1), "sumsq": (d * d).sum(axis=1),
            "full_b": np.array(fb, dtype=np.int64), "y_full": flat[fb].copy()}


def run_case(name: str) -> Path:
    from . import ref_import
    overrides, state, batch, frames, seed = CONFIG_PINS[name]
    models, _ = ref_import.load_reference()
    h = ref_import.RefHParams(**overrides)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gen = fixtures.construct(models.Generator, h)
        fixtures.apply_state(gen, state)
    x, s, n = fixtures.make_inputs(h, batch, frames, seed)
    t0 = time.time()
    with torch.no_grad():
        y = gen(x, s, n).numpy()
    dt = time.time() - t0
    arrays = summarise(y)
    meta = dict(name=name, overrides={k: (v if not isinstance(v, str) else f"str:{v}") for k, v in overrides.items()},
                state=state, batch=batch, frames=frames, seed=seed, torch=torch.__version__,
                block=type(gen.resblocks[0]).__name__, stride=PIN_STRIDE,
                abs_max=float(np.abs(y).max()), rms=float(np.sqrt((y.astype(np.float64) ** 2).mean())))
    arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    for k, v in fixtures.small_state(gen).items():
        arrays["small/" + k] = v
    OUT_DIR.mkdir(parents=True, exist_ok=True)
    path = OUT_DIR / f"{name}.npz"
    np.savez_compressed(path, **arrays)
    print(f"{name}: {meta['block']} B={batch} T={frames} C={overrides['num_wv_feat']} reference forward {dt:.1f} s "
          f"|y|max={meta['abs_max']:.4f} rms={meta['rms']:.4f} pins={arrays['idx'].size} "
          f"-> {path.name} ({path.stat().st_size / 1024:.0f} KiB)", flush=True)
    return path


def main(argv):
    for n in (argv or list(CONFIG_PINS)):
        run_case(n)


if __name__ == "__main__":
    main(sys.argv[1:])
