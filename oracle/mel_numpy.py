"""numpy restatement of the reference's log-mel spectrogram (TEST INFRASTRUCTURE ONLY).

Follows ``mel_spectrogram`` of vec2wav/dataset.py:53-77 (called at vec2wav/train.py:172-173,266-268,
282-284 with vec2wav/hparams.py:50-61): reflect-pad by (n_fft - hop)/2 (:66-68), torch.stft with a
periodic Hann window, center=False, one-sided (:70-71), sqrt(re^2 + im^2 + 1e-9) (:73), the
librosa.filters.mel filterbank (:62) and log(clamp(., 1e-5)) (:76, :33-34).

The filterbank lives in a third-party dependency that is absent from /root/reference AND from this image:
librosa (requirements.txt pins no version; `librosa_mel_fn(sampling_rate, n_fft, num_mels, fmin, fmax)`
is called positionally, i.e. librosa < 0.10).  `slaney_mel_filterbank` restates its published algorithm
(librosa/filters.py `mel`, defaults htk=False, norm='slaney'; librosa/core/convert.py `hz_to_mel`,
`mel_to_hz`, `mel_frequencies`) in float64 with a final cast to float32 as librosa does.  It is pinned
against torchaudio.functional.melscale_fbanks(norm='slaney', mel_scale='slaney') -- an independent
implementation of the same definition -- and everything else against the reference function itself run
in the build container (oracle/make_golden.py:run_mel_case, librosa import shimmed with this filterbank).
"""
from __future__ import annotations

import numpy as np


def hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mels)


def mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def slaney_mel_filterbank(sr, n_fft, n_mels=128, fmin=0.0, fmax=None) -> np.ndarray:
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax): [n_mels, 1 + n_fft//2] float32."""
    if fmax is None:
        fmax = float(sr) / 2
    weights = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float64)
    fftfreqs = np.linspace(0, float(sr) / 2, 1 + n_fft // 2)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights.astype(np.float32)


def mel_spectrogram(y, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, dtype=np.float64):
    """y [B, L] -> log-mel [B, num_mels, frames]   (vec2wav/dataset.py:53-77, center=False)."""
    y = np.asarray(y, dtype=dtype)
    pad = int((n_fft - hop_size) / 2)
    yp = np.pad(y, ((0, 0), (pad, pad)), mode="reflect")                       # :66-68
    frames = (yp.shape[1] - n_fft) // hop_size + 1
    idx = np.arange(n_fft)[None, :] + hop_size * np.arange(frames)[:, None]
    win = np.zeros(n_fft, dtype=dtype)
    left = (n_fft - win_size) // 2
    win[left:left + win_size] = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(win_size) / win_size)   # torch.hann_window
    spec = np.fft.rfft(yp[:, idx] * win, axis=-1)                              # [B, frames, n_fft/2+1]
    mag = np.sqrt(spec.real ** 2 + spec.imag ** 2 + 1e-9)                      # :73
    fb = slaney_mel_filterbank(sampling_rate, n_fft, num_mels, fmin, fmax).astype(dtype)
    mel = np.einsum("mk,bfk->bmf", fb, mag)                                    # :75
    return np.log(np.clip(mel, 1e-5, None))                                    # :76
