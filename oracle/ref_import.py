"""Import the UNMODIFIED reference Generator from /root/reference (build container only).

`vec2wav/utils.py:3,7-8` imports matplotlib (absent here) for a plotting helper the Generator
never calls; two stub modules satisfy the import.  The GPU box has no /root/reference: anything
that runs there must use the committed fixtures instead (tests/golden/).
"""
from __future__ import annotations

import importlib
import sys
import types
import warnings
from pathlib import Path

REFERENCE_ROOT = Path("/root/reference")


def available() -> bool:
    return (REFERENCE_ROOT / "vec2wav" / "models.py").exists()


def load_reference():
    """Returns (models_module, hparams_module) of the reference's vec2wav package."""
    if not available():
        raise RuntimeError("/root/reference is not present on this machine")
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        mpl.use = lambda *a, **k: None
        pylab = types.ModuleType("matplotlib.pylab")
        mpl.pylab = pylab
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pylab"] = pylab
    path = str(REFERENCE_ROOT / "vec2wav")
    if path not in sys.path:
        sys.path.insert(0, path)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        hparams = importlib.import_module("hparams")
        models = importlib.import_module("models")
    return models, hparams


class RefHParams:
    """Attribute bag with the reference's generator hyper-parameters and overrides."""

    def __init__(self, **overrides):
        _, hp = load_reference()
        for name in ("num_wv_feat", "upsample_initial_channel", "upsample_rates",
                     "upsample_kernel_sizes", "resblock", "resblock_kernel_sizes",
                     "resblock_dilation_sizes", "spk_dim", "noise_dim", "seed"):
            setattr(self, name, getattr(hp, name))
        for k, v in overrides.items():
            setattr(self, k, v)
