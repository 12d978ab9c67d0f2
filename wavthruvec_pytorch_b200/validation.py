"""The reference's validation step on the native path (caller side).

vec2wav/train.py:246-288 puts the generator in eval(), generates every validation item under
no_grad, takes the log-mel of the generated waveform (dataset.py:53-77) and accumulates the L1
distance to the ground-truth mel (`validation/mel_spec_error`).  Here the two heavy steps are libv2w
calls (v2w_forward, v2w_mel_forward); the permute / truncate / mean of train.py:269-274 is a handful of
elementwise torch ops on a [B, frames, 80] tensor.
"""
from __future__ import annotations

from typing import Iterable, Optional, Tuple

import torch

from .mel import mel_spectrogram

#: vec2wav/hparams.py:50-61 (fmax_for_loss = None is what train.py:266-268 passes)
MEL_DEFAULTS = dict(n_fft=1024, num_mels=80, sampling_rate=16000, hop_size=256, win_size=1024, fmin=0,
                    fmax_for_loss=None)


def _mel_args(hp) -> Tuple:
    get = (lambda k: getattr(hp, k, MEL_DEFAULTS[k])) if hp is not None else (lambda k: MEL_DEFAULTS[k])
    return (get("n_fft"), get("num_mels"), get("sampling_rate"), get("hop_size"), get("win_size"),
            get("fmin"), get("fmax_for_loss"))


def mel_spec_error(generator, wv_feat: torch.Tensor, spk_emb: torch.Tensor, y_mel: torch.Tensor,
                   noise: Optional[torch.Tensor] = None, hp=None,
                   mel_fn=mel_spectrogram) -> Tuple[float, torch.Tensor]:
    """One validation item (train.py:253-274): returns (L1 mel error, generated waveform [B, 1, L]).

    wv_feat [B, C, T], spk_emb [B, spk_dim], y_mel [B, frames_gt, num_mels] (the dataset's layout after
    collate); the ground truth is truncated to the generated frame count as train.py:273 does.  The
    caller is responsible for eval() mode, exactly like the reference loop."""
    if generator.training:
        raise RuntimeError("mel_spec_error is the eval()-mode validation step (train.py:246)")
    with torch.no_grad():
        if noise is None:
            noise = torch.randn(wv_feat.shape[0], generator.h.noise_dim, device=wv_feat.device)   # train.py:259
        y_g_hat = generator(wv_feat, spk_emb, noise)                                               # :262
        y_g_hat_mel = mel_fn(y_g_hat.squeeze(1), *_mel_args(hp)).permute(0, 2, 1)                  # :266-269
        y_mel = y_mel.to(y_g_hat_mel.device)[:, :y_g_hat_mel.shape[1], :]                          # :273
        if y_mel.shape != y_g_hat_mel.shape:
            raise RuntimeError(f"ground-truth mel {tuple(y_mel.shape)} does not cover the generated "
                               f"{tuple(y_g_hat_mel.shape)}")
        err = torch.nn.functional.l1_loss(y_mel, y_g_hat_mel).item()                               # :274
    return err, y_g_hat


def validate(generator, batches: Iterable, hp=None) -> float:
    """`validation/mel_spec_error` over an iterable of (wv_feat, spk_emb, y_mel) device batches
    (train.py:251-288): mean of the per-batch L1 errors; restores train() mode if it was set."""
    was_training = generator.training
    generator.eval()
    total, n = 0.0, 0
    try:
        for wv_feat, spk_emb, y_mel in batches:
            err, _ = mel_spec_error(generator, wv_feat, spk_emb, y_mel, hp=hp)
            total += err
            n += 1
    finally:
        if was_training:
            generator.train()
    if n == 0:
        raise ValueError("no validation batches")
    return total / n
