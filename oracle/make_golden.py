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


MEL_PARAMS = dict(n_fft=1024, num_mels=80, sampling_rate=16000, hop_size=256, win_size=1024, fmin=0, fmax=None)
#                 vec2wav/hparams.py:50-61 (fmax_for_loss = None is what train.py:172-173,266-268 pass)


def mel_test_waves(batch: int, samples: int, seed: int) -> np.ndarray:
    """Seeded speech-like test signals in (-1, 1): a few drifting harmonics, noise bursts, a silent gap."""
    rng = np.random.default_rng(seed)
    t = np.arange(samples) / 16000.0
    out = np.zeros((batch, samples), dtype=np.float64)
    for b in range(batch):
        f0 = rng.uniform(90, 260)
        for hnum in range(1, 12):
            out[b] += rng.uniform(0.02, 0.2) / hnum * np.sin(2 * np.pi * f0 * hnum * t * (1 + 0.05 * np.sin(2 * np.pi * 3 * t))
                                                             + rng.uniform(0, 6.28))
        out[b] += 0.02 * rng.standard_normal(samples) * (np.sin(2 * np.pi * 2.5 * t + b) > 0.3)
        gap = rng.integers(0, max(1, samples - samples // 8))
        out[b, gap: gap + samples // 8] *= 1e-4
    return np.clip(out, -0.99, 0.99).astype(np.float32)


def run_mel_case() -> Path:
    """Log-mel of seeded waveforms by the reference's own mel_spectrogram (vec2wav/dataset.py:53-77), run
    unmodified; librosa is absent from this image, so its import is shimmed: librosa.filters.mel ->
    oracle.mel_numpy.slaney_mel_filterbank (restated from librosa's published algorithm; the file also
    stores torchaudio's independent slaney filterbank for the cross-check)."""
    import importlib
    import types
    import torchaudio
    from . import mel_numpy
    ref_import.load_reference()
    lib, util, filt = (types.ModuleType(n) for n in ("librosa", "librosa.util", "librosa.filters"))
    util.normalize = lambda x, *a, **k: x
    filt.mel = mel_numpy.slaney_mel_filterbank
    lib.util, lib.filters = util, filt
    sys.modules.update({"librosa": lib, "librosa.util": util, "librosa.filters": filt})
    ds = importlib.import_module("dataset")
    arrays = {}
    for tag, (batch, samples, seed) in {"a": (2, 16000, 31), "b": (3, 5000, 32)}.items():
        y = mel_test_waves(batch, samples, seed)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ds.mel_basis.clear(); ds.hann_window.clear()
            m = ds.mel_spectrogram(torch.from_numpy(y), **MEL_PARAMS)
        arrays[f"y_{tag}"] = y
        arrays[f"mel_{tag}"] = m.numpy()
    arrays["fb_torchaudio"] = torchaudio.functional.melscale_fbanks(
        513, 0.0, 8000.0, 80, 16000, norm="slaney", mel_scale="slaney").T.contiguous().numpy()
    path = OUT_DIR / "mel_reference.npz"
    np.savez_compressed(path, **arrays)
    print(f"mel_reference: {[(k, v.shape) for k, v in arrays.items()]} -> {path.name} ({path.stat().st_size/1024:.0f} KiB)")
    return path


TRAIN_CASES = ("rb2_768_stress", "rb1_1024_stress")


def main(argv):
    names = argv or list(fixtures.GOLDEN_CASES)
    for n in names:
        if n == "mel":
            run_mel_case()
        elif n.endswith("_train"):
            run_train_case(n[:-len("_train")])
        else:
            run_case(n)
    if not argv:
        for n in TRAIN_CASES:
            run_train_case(n)
        run_mel_case()


if __name__ == "__main__":
    main(sys.argv[1:])
