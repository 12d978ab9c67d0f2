"""Log-mel spectrogram of generated waveforms on the device.

Drop-in for ``mel_spectrogram`` of the reference's vec2wav/dataset.py:53-77 -- the function
vec2wav/train.py:172-173,266-268,282-284 runs on every generator output for the mel loss and in
validation -- with the same signature.  All arithmetic happens in libv2w (``v2w_mel_*``: a framing
kernel, the STFT as one tcgen05 tap-GEMM with split-fp16 operands, a magnitude + filterbank + log
kernel); there is no CPU path.

Autograd: the native kernels have no backward.  When gradients are required (the mel-L1 loss of
train.py:172-174 with autograd enabled) the call is delegated to the reference's own differentiable op
sequence (reflect pad -> torch.stft -> magnitude -> filterbank matmul -> log-clamp) with the library's
filterbank, exactly like ``Generator`` delegates its autograd forward -- a mode delegation, never silent:
the native path is taken only under ``torch.no_grad()`` or for inputs that do not require grad.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Tuple

import numpy as np
import torch

from . import native


class _MelPlan:
    def __init__(self, lib, cfg: native.MelConfig, batch: int, samples: int, device: torch.device):
        self._lib = lib
        handle = C.c_void_p()
        native.check(lib.v2w_mel_create(C.byref(cfg), batch, samples, C.byref(handle)))
        self.handle = handle
        self.frames = int(lib.v2w_mel_num_frames(handle))
        self.num_mels = int(cfg.num_mels)
        self.nbins = int(cfg.n_fft) // 2 + 1
        self.ws = None
        if device.type == "cuda":
            self.ws = torch.empty(int(lib.v2w_mel_workspace_bytes(handle)), dtype=torch.uint8, device=device)
            with torch.cuda.device(device):
                native.check(lib.v2w_mel_bind(handle, self.ws.data_ptr()))

    def filterbank(self) -> np.ndarray:
        fb = np.empty((self.num_mels, self.nbins), dtype=np.float32)
        native.check(self._lib.v2w_mel_get_filterbank(self.handle, fb.ctypes.data))
        return fb

    def __del__(self):
        try:
            if self.handle:
                self._lib.v2w_mel_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


_PLANS: Dict[Tuple, _MelPlan] = {}
last_launches = 0


def _config(n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax) -> native.MelConfig:
    return native.MelConfig(int(n_fft), int(hop_size), int(win_size), int(num_mels), int(sampling_rate),
                            float(fmin), float(fmax) if fmax is not None else 0.0)


def filterbank(n_fft, num_mels, sampling_rate, fmin, fmax, hop_size=None, win_size=None) -> np.ndarray:
    """The library's Slaney filterbank [num_mels, n_fft/2+1] (host only; = librosa.filters.mel defaults)."""
    hop = int(hop_size or n_fft // 4)
    cfg = _config(n_fft, num_mels, sampling_rate, hop, win_size or n_fft, fmin, fmax)
    return _MelPlan(native.load(), cfg, 1, 4 * int(n_fft), torch.device("cpu")).filterbank()


_AUTOGRAD_CONSTS: Dict[Tuple, Tuple[torch.Tensor, torch.Tensor]] = {}


def _mel_autograd(y, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax) -> torch.Tensor:
    """Differentiable delegate: the op sequence of vec2wav/dataset.py:53-77 on stock PyTorch ops."""
    key = (int(n_fft), int(num_mels), int(sampling_rate), int(hop_size), int(win_size), float(fmin),
           None if fmax is None else float(fmax), y.device)
    consts = _AUTOGRAD_CONSTS.get(key)
    if consts is None:
        fb = torch.from_numpy(filterbank(n_fft, num_mels, sampling_rate, fmin, fmax, hop_size, win_size)).to(y.device)
        consts = (fb, torch.hann_window(int(win_size), device=y.device))
        _AUTOGRAD_CONSTS[key] = consts
    fb, window = consts
    pad = int((n_fft - hop_size) / 2)
    yp = torch.nn.functional.pad(y.unsqueeze(1), (pad, pad), mode="reflect").squeeze(1)
    spec = torch.stft(yp, int(n_fft), hop_length=int(hop_size), win_length=int(win_size), window=window,
                      center=False, pad_mode="reflect", normalized=False, onesided=True, return_complex=True)
    mag = torch.sqrt(spec.real.pow(2) + spec.imag.pow(2) + 1e-9)
    return torch.log(torch.clamp(torch.matmul(fb, mag), min=1e-5))


def mel_spectrogram(y: torch.Tensor, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax,
                    center: bool = False) -> torch.Tensor:
    """y [B, L] in [-1, 1] (CUDA) -> log-mel [B, num_mels, L // hop_size]  (vec2wav/dataset.py:53-77).

    Native (non-differentiable) kernels under ``no_grad`` / for inputs without grad; the differentiable
    PyTorch delegate when ``y.requires_grad`` and autograd is enabled (see the module docstring)."""
    global last_launches
    if center:
        raise NotImplementedError("center=True is not supported (the reference always passes center=False)")
    if y.dim() != 2:
        raise RuntimeError(f"expected y of shape [B, L], got {tuple(y.shape)}")
    if torch.is_grad_enabled() and y.requires_grad:
        last_launches = 0
        return _mel_autograd(y.to(torch.float32), n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax)
    if not y.is_cuda:
        raise RuntimeError("the native mel-spectrogram runs on sm_100 GPUs only and has no CPU fallback")
    y = y.detach().to(torch.float32).contiguous()
    B, L = y.shape
    key = (B, L, int(n_fft), int(num_mels), int(sampling_rate), int(hop_size), int(win_size), float(fmin),
           None if fmax is None else float(fmax), y.device)
    plan = _PLANS.get(key)
    if plan is None:
        if len(_PLANS) > 64:
            _PLANS.clear()
        plan = _MelPlan(native.load(), _config(n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax),
                        B, L, y.device)
        _PLANS[key] = plan
    out = torch.empty((B, plan.num_mels, plan.frames), dtype=torch.float32, device=y.device)
    n = C.c_int32(0)
    with torch.cuda.device(y.device):
        stream = torch.cuda.current_stream(y.device).cuda_stream
        native.check(native.load().v2w_mel_forward(plan.handle, y.data_ptr(), out.data_ptr(),
                                                   C.c_void_p(stream), C.byref(n)))
    last_launches = n.value
    return out
