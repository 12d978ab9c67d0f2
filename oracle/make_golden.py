"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

    python -m oracle.make_golden            # all cases of oracle.fixtures.GOLDEN_CASES

Each file holds: the inputs, the reference's fp32 output (post-tanh) and conv_post output
(pre-tanh, via a forward hook), a float64 re-run of the same module for the noise floor,
the warm-up products (BN running stats, spectral-norm u/v) and per-tensor checksums of the whole
state_dict so that a differently-initialised module is detected.
"""
from __future__ import annotations

import json
import sys
import warnings
from pathlib import Path

import numpy as np
import torch

from . import fixtures, ref_import

OUT_DIR = Path(__file__).resolve().parent.parent / "tests" / "golden"


def run_case(name: str) -> Path:
    overrides, state, batch, frames, seed = fixtures.GOLDEN_CASES[name]
    models, _ = ref_import.load_reference()
    h = ref_import.RefHParams(**overrides)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gen = fixtures.construct(models.Generator, h)
        fixtures.apply_state(gen, state)
    x, s, n = fixtures.make_inputs(h, batch, frames, seed)
    pre = {}
    hook = gen.conv_post.register_forward_hook(lambda m, i, o: pre.__setitem__("y", o.detach().clone()))
    with torch.no_grad():
        y = gen(x, s, n)
    hook.remove()
    with torch.no_grad():
        gen64 = gen.double()
        y64 = gen64(x.double(), s.double(), n.double())
        gen.float()
    small = fixtures.small_state(gen)
    sums = fixtures.state_checksums(gen)
    meta = dict(name=name, overrides={k: (v if not isinstance(v, str) else f"str:{v}")
                                      for k, v in overrides.items()},
                state=state, batch=batch, frames=frames, seed=seed,
                torch=torch.__version__, block=type(gen.resblocks[0]).__name__,
                num_keys=len(gen.state_dict()),
                num_params=sum(p.numel() for p in gen.parameters()))
    arrays = {
        "x": x.numpy(), "spk": s.numpy(), "noise": n.numpy(),
        "y": y.numpy(), "y_pre": pre["y"].numpy(), "y64": y64.numpy(),
        "meta": np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8),
        "checksum_keys": np.array(list(sums.keys())),
        "checksum_vals": np.array(list(sums.values()), dtype=np.float64),
    }
    for k, v in small.items():
        arrays["small/" + k] = v
    OUT_DIR.mkdir(parents=True, exist_ok=True)
    path = OUT_DIR / f"{name}.npz"
    np.savez_compressed(path, **arrays)
    sat = float((np.abs(y.numpy()) > 0.99).mean())
    print(f"{name}: block={meta['block']} keys={meta['num_keys']} params={meta['num_params']} "
          f"|y|max={np.abs(y.numpy()).max():.4f} rms={np.sqrt((y.numpy()**2).mean()):.4f} "
          f"sat={sat:.3f} fp32-vs-fp64 maxabs={np.abs(y.numpy()-y64.numpy()).max():.2e} "
          f"-> {path.name} ({path.stat().st_size/1024:.0f} KiB)")
    return path


TRAIN_STEPS = 2


def run_train_case(name: str) -> Path:
    """train()-mode forwards under no_grad: from the fixture state of `name`,
    TRAIN_STEPS forwards of the golden inputs with module.training set; stores every output and the
    buffers the forwards mutate (BN running stats + counter, spectral-norm u / v) after each step."""
    overrides, state, batch, frames, seed = fixtures.GOLDEN_CASES[name]
    models, _ = ref_import.load_reference()
    h = ref_import.RefHParams(**overrides)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gen = fixtures.construct(models.Generator, h)
        fixtures.apply_state(gen, state)
    x, s, n = fixtures.make_inputs(h, batch, frames, seed)
    arrays = {"steps": np.array(TRAIN_STEPS)}
    for k, v in fixtures.small_state(gen).items():
        arrays["small/" + k] = v
    gen.train()
    for step in range(TRAIN_STEPS):
        with torch.no_grad():
            y = gen(x * (1.0 + 0.25 * step), s, n)
        arrays[f"y{step}"] = y.numpy()
        for k, v in fixtures.small_state(gen).items():
            arrays[f"after{step}/" + k] = v
    path = OUT_DIR / f"{name}_train.npz"
    np.savez_compressed(path, **arrays)
    print(f"{name}_train: {TRAIN_STEPS} train() forwards, |y|max={np.abs(y.numpy()).max():.4f} "
          f"-> {path.name} ({path.stat().st_size/1024:.0f} KiB)")
    return path


TRAIN_CASES = ("rb2_768_stress", "rb1_1024_stress")


def main(argv):
    names = argv or list(fixtures.GOLDEN_CASES)
    for n in names:
        if n.endswith("_train"):
            run_train_case(n[:-len("_train")])
        else:
            run_case(n)
    if not argv:
        for n in TRAIN_CASES:
            run_train_case(n)


if __name__ == "__main__":
    main(sys.argv[1:])
