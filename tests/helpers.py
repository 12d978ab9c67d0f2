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
