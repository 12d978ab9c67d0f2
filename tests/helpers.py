"""Shared test helpers: golden loading and fixture-state construction of the drop-in module."""
from __future__ import annotations

import json
import warnings
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

from oracle import fixtures

GOLDEN_DIR = Path(__file__).resolve().parent / "golden"


def hparams(**overrides):
    from wavthruvec_pytorch_b200.hparams import HParams
    h = HParams()
    for k, v in overrides.items():
        setattr(h, k, v)
    return h


def load_golden(name: str):
    z = np.load(GOLDEN_DIR / f"{name}.npz")
    meta = json.loads(bytes(z["meta"]).decode())
    overrides = {k: (v[4:] if isinstance(v, str) and v.startswith("str:") else v)
                 for k, v in meta["overrides"].items()}
    small = {k[len("small/"):]: z[k] for k in z.files if k.startswith("small/")}
    sums = dict(zip([str(k) for k in z["checksum_keys"]], z["checksum_vals"]))
    return SimpleNamespace(name=name, meta=meta, overrides=overrides, state=meta["state"],
                           x=z["x"], spk=z["spk"], noise=z["noise"], y=z["y"], y_pre=z["y_pre"],
                           y64=z["y64"], small=small, checksums=sums)


def load_train_golden(name: str):
    """tests/golden/<name>_train.npz (oracle/make_golden.py:run_train_case): train()-mode forwards of
    the reference from the fixture state of golden `name`; step k feeds x * (1 + 0.25 k)."""
    z = np.load(GOLDEN_DIR / f"{name}_train.npz")
    steps = int(z["steps"])
    return SimpleNamespace(
        steps=steps, y=[z[f"y{k}"] for k in range(steps)],
        after=[{key[len(f"after{k}/"):]: z[key] for key in z.files if key.startswith(f"after{k}/")}
               for k in range(steps)])


_MODULE_CACHE = {}


def dropin_in_state(g, cache: bool = True):
    """The drop-in Generator in the golden's fixture state, built WITHOUT the reference and
    without warm-up forwards (seeded init + seeded re-draw + committed warm-up products)."""
    if cache and g.name in _MODULE_CACHE:
        return _MODULE_CACHE[g.name]
    from wavthruvec_pytorch_b200 import Generator
    h = hparams(**g.overrides)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gen = fixtures.construct(Generator, h)
    fixtures.apply_state_from_golden(gen, g.state, g.small)
    if cache:
        _MODULE_CACHE[g.name] = gen
    return gen


def assert_checksums(gen, g, rtol=0.0):
    got = fixtures.state_checksums(gen)
    assert set(got) == set(g.checksums), "state_dict key sets differ"
    for k, (s1, s2) in got.items():
        e1, e2 = g.checksums[k]
        assert s1 == e1 and s2 == e2, f"{k}: checksum ({s1},{s2}) != golden ({e1},{e2})"


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


# ----------------------------------------------------------------------------------------------
# config-size pins (oracle/make_config_pins.py): compact pins of reference outputs at BASELINE sizes
# ----------------------------------------------------------------------------------------------
def load_config_pin(name: str):
    z = np.load(GOLDEN_DIR / f"{name}.npz")
    meta = json.loads(bytes(z["meta"]).decode())
    overrides = {k: (v[4:] if isinstance(v, str) and v.startswith("str:") else v)
                 for k, v in meta["overrides"].items()}
    small = {k[len("small/"):]: z[k] for k in z.files if k.startswith("small/")}
    return SimpleNamespace(name=name, meta=meta, overrides=overrides, state=meta["state"],
                           batch=meta["batch"], frames=meta["frames"], seed=meta["seed"], small=small,
                           idx=z["idx"], y_s=z["y_s"], sum=z["sum"], sumsq=z["sumsq"],
                           full_b=z["full_b"], y_full=z["y_full"])


def pin_inputs(pin):
    """The seeded inputs of a config-size pin (regenerated, never stored)."""
    h = hparams(**pin.overrides)
    return fixtures.make_inputs(h, pin.batch, pin.frames, pin.seed)


def compare_with_pin(pin, y) -> dict:
    """y: [B, 1, L] fp32 waveform (numpy) produced from pin_inputs(pin).  Compares the strided samples,
    the whole stored utterances and the per-utterance float64 sums with what the reference produced."""
    flat = np.ascontiguousarray(y, dtype=np.float32).reshape(pin.batch, -1)
    got_s = flat.reshape(-1)[pin.idx]
    got_f = flat[pin.full_b]
    ref = np.concatenate([pin.y_s.reshape(-1), pin.y_full.reshape(-1)]).astype(np.float64)
    got = np.concatenate([got_s.reshape(-1), got_f.reshape(-1)]).astype(np.float64)
    d = flat.astype(np.float64)
    energy = np.sqrt(np.maximum(pin.sumsq, 1e-30))
    return {
        "max_abs": float(np.abs(ref - got).max()),
        "snr_db": fixtures.snr_db(ref, got),
        # per-utterance sums cover every sample of every utterance: |sum(y) - sum(ref)| / sqrt(L sumsq)
        "sum_err": float((np.abs(d.sum(axis=1) - pin.sum) / (energy * np.sqrt(flat.shape[1]))).max()),
        "sumsq_rel": float((np.abs((d * d).sum(axis=1) - pin.sumsq) / np.maximum(pin.sumsq, 1e-30)).max()),
        "nan": bool(np.isnan(flat).any()),
    }


def dropin_for_pin(pin):
    from wavthruvec_pytorch_b200 import Generator
    h = hparams(**pin.overrides)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gen = fixtures.construct(Generator, h)
    fixtures.apply_state_from_golden(gen, pin.state, pin.small)
    return gen
