"""Generator hyper-parameters (constructor contract of the hot path).

The reference passes its *module object* ``hparams`` as ``h`` (vec2wav/train.py:22,65);
``Generator(h)`` only ever reads the attributes listed below (vec2wav/models.py:81-110).
``vec2wav/config.py`` in the reference is an empty file; the values live in
vec2wav/hparams.py:24-57 and are mirrored here as plain data so that a user can write
``Generator(HParams())`` or keep passing the reference's own ``hparams`` module.

Only generator fields are carried: discriminator / training / path constants are out of
scope for this package.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List


@dataclass
class HParams:
    # vec2wav/hparams.py:26-27
    spk_dim: int = 192
    noise_dim: int = 192
    # vec2wav/hparams.py:30  -- an *int* in the shipped file; the selector in
    # vec2wav/models.py:84 compares against the string '1', so the shipped default
    # builds ResBlock2.  Pass resblock='1' for ResBlock1.
    resblock: object = 1
    # vec2wav/hparams.py:40-44
    upsample_rates: List[int] = field(default_factory=lambda: [5, 4, 4, 2, 2])
    upsample_kernel_sizes: List[int] = field(default_factory=lambda: [11, 8, 8, 4, 4])
    upsample_initial_channel: int = 512
    resblock_kernel_sizes: List[int] = field(default_factory=lambda: [3, 7, 11])
    resblock_dilation_sizes: List[List[int]] = field(
        default_factory=lambda: [[1, 3, 5], [1, 3, 5], [1, 3, 5]])
    # vec2wav/hparams.py:51 (1024 = huggingface wav2vec 2.0 large; 768 = base)
    num_wv_feat: int = 1024
    # vec2wav/hparams.py:37,57 -- not read by Generator, kept because callers use them
    seed: int = 1234
    sampling_rate: int = 16000


#: frames per second of wav2vec 2.0 features (20 ms stride, readme.md:17,37-39)
FRAMES_PER_SECOND = 50
#: product of upsample_rates: 16 kHz / 50 Hz
SAMPLES_PER_FRAME = 320
