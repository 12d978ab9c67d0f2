"""Drop-in for ``Generator`` of the reference's vec2wav/models.py (77-156).

Same constructor (`Generator(h)` with the attribute bag of vec2wav/hparams.py), same
``forward(x, spk_emb, noise)``, same ``state_dict`` keys / shapes / dtypes (156 keys with
ResBlock2, 336 with ResBlock1), same ``remove_weight_norm()``.  What differs is where the
arithmetic runs:

* ``eval()`` under ``torch.no_grad()`` on a CUDA device -- the generation path -- executes in
  libv2w (hand-written sm_100a kernels behind the C-ABI of include/v2w.h).  Weight norm,
  spectral norm and BN running statistics are folded once per parameter version instead of on
  every call.  There is no CPU or eager fallback for this path: without the library or without
  an sm_100 GPU it raises.
* ``train()`` under ``torch.no_grad()`` on a CUDA device (validation and warm-up loops) also runs
  in libv2w (``v2w_forward_train``): batch-statistic BN, one spectral-norm power iteration per
  stage and the in-place buffer updates of the reference.
* With grad enabled the forward needs autograd; it runs through stock PyTorch ops with the
  reference's exact op order (``_forward_autograd``; also ``train()`` on CPU tensors).  That is a
  mode delegation, not a fallback of the native path.
"""
from __future__ import annotations

import os
import warnings
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.utils import remove_weight_norm as _strip_wn
from torch.nn.utils import weight_norm as _wn

from .modules import Z_CHANNELS, ConditionalBatchNorm1d
from .utils import get_padding, init_weights

LRELU_SLOPE = 0.1  # vec2wav/models.py:10


def _wn_conv(channels: int, k: int, dilation: int) -> nn.Module:
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", FutureWarning)
        return _wn(nn.Conv1d(channels, channels, k, 1, dilation=dilation,
                             padding=get_padding(k, dilation)))


class ResBlock2(nn.Module):
    """Two dilated convs, each ``x = conv(lrelu(x)) + x`` (vec2wav/models.py:53-74).

    Keys: ``convs.{0,1}.{bias,weight_g,weight_v}``; only dilation[0:2] is used.
    """

    def __init__(self, h, channels: int, kernel_size: int = 3, dilation=(1, 3)):
        super().__init__()
        self.h = h
        self.convs = nn.ModuleList(_wn_conv(channels, kernel_size, dilation[s]) for s in range(2))
        self.convs.apply(init_weights)

    def exec_order(self):
        return list(self.convs)

    def forward(self, x):
        for conv in self.convs:
            x = conv(F.leaky_relu(x, LRELU_SLOPE)) + x
        return x

    def remove_weight_norm(self):
        for conv in self.convs:
            _strip_wn(conv)


class ResBlock1(nn.Module):
    """Three (dilated conv, plain conv) pairs with a residual around each pair
    (vec2wav/models.py:13-50).  Keys: ``convs1.{0,1,2}.*``, ``convs2.{0,1,2}.*``."""

    def __init__(self, h, channels: int, kernel_size: int = 3, dilation=(1, 3, 5)):
        super().__init__()
        self.h = h
        self.convs1 = nn.ModuleList(_wn_conv(channels, kernel_size, dilation[s]) for s in range(3))
        self.convs1.apply(init_weights)
        self.convs2 = nn.ModuleList(_wn_conv(channels, kernel_size, 1) for _ in range(3))
        self.convs2.apply(init_weights)

    def exec_order(self):
        out = []
        for a, b in zip(self.convs1, self.convs2):
            out += [a, b]
        return out

    def forward(self, x):
        for a, b in zip(self.convs1, self.convs2):
            y = a(F.leaky_relu(x, LRELU_SLOPE))
            y = b(F.leaky_relu(y, LRELU_SLOPE))
            x = y + x
        return x

    def remove_weight_norm(self):
        for conv in list(self.convs1) + list(self.convs2):
            _strip_wn(conv)


class Generator(nn.Module):
    def __init__(self, h, precision: Optional[str] = None):
        super().__init__()
        self.h = h
        self.num_kernels = len(h.resblock_kernel_sizes)
        self.num_upsamples = len(h.upsample_rates)
        c0 = h.upsample_initial_channel
        # --- registration order == reference (RNG consumption and state_dict order) -----------
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", FutureWarning)
            self.conv_pre = _wn(nn.Conv1d(h.num_wv_feat, c0, 7, 1, padding=3))
            block_cls = ResBlock1 if h.resblock == '1' else ResBlock2  # models.py:84 (string compare)
            self.ups = nn.ModuleList()
            for i, (u, k) in enumerate(zip(h.upsample_rates, h.upsample_kernel_sizes)):
                self.ups.append(_wn(nn.ConvTranspose1d(c0 >> i, c0 >> (i + 1), k, u,
                                                       padding=(k - u) // 2)))
            self.resblocks = nn.ModuleList()
            ch = c0
            for i in range(len(self.ups)):
                ch = c0 >> (i + 1)
                for k, d in zip(h.resblock_kernel_sizes, h.resblock_dilation_sizes):
                    self.resblocks.append(block_cls(h, ch, k, d))
            self.conv_post = _wn(nn.Conv1d(ch, 1, 7, 1, padding=3))
        self.ups.apply(init_weights)
        self.conv_post.apply(init_weights)
        self.cbns = nn.ModuleList()
        self.fcs = nn.ModuleList()
        for i in range(len(self.ups)):
            self.fcs.append(nn.Linear(h.spk_dim + h.noise_dim, Z_CHANNELS))
            self.cbns.append(ConditionalBatchNorm1d(256 >> i))  # hard-coded widths: models.py:113
        # --- native engine state (not part of the state_dict) ---------------------------------
        object.__setattr__(self, "_engine", None)
        self._precision = precision
        self._warned_grad = False
        #: replay eval()-mode forwards from a CUDA graph (latency path for batch-1 / long-form requests;
        #: see v2w_forward_graph in include/v2w.h).  Also via $V2W_CUDA_GRAPH=1.
        self.use_cuda_graph = bool(int(os.environ.get("V2W_CUDA_GRAPH", "0") or 0))
        #: serving: set True once the checkpoint is loaded and the module will not change any more -- the native
        #: path then skips its per-call change detection (parameter versions / storage pointers).  Any later
        #: parameter edit must be followed by `invalidate()` (or by setting this back to False).
        self.parameters_frozen = False
        #: eval()-mode batches larger than this many frames (batch x frames) run as consecutive micro-batches
        #: through one plan (utterances are independent; see NativeEngine.forward).  None = never split.
        self.max_batch_frames = int(os.environ.get("V2W_MAX_BATCH_FRAMES", "32768") or 0) or None

    # ------------------------------------------------------------------------------------------
    # native engine plumbing
    # ------------------------------------------------------------------------------------------
    @property
    def precision(self) -> str:
        """Arithmetic of the native path: 'bf16' / 'fp16' (tcgen05 with 16-bit operands, fp32
        accumulation and fp32 residual stream), 'fp32tc' (fp32 values on the tensor cores: split-fp16
        operands, three MMAs per product, meets the fp32 bar) or 'fp32' (FFMA).  Default 'auto' (also via
        $V2W_PRECISION): bf16 for ResBlock2; fp16 for ResBlock1, whose 97-conv chain does not reach
        the 45 dB waveform-SNR bar with bf16 operands."""
        import os
        value = (self._precision or os.environ.get("V2W_PRECISION", "auto")).lower()
        if value == "auto":
            value = "fp16" if self.h.resblock == '1' else "bf16"
        return value

    @property
    def train_precision(self) -> str:
        """Arithmetic of the native train()-mode forward.  'auto' means fp16 operands for both block
        types: batch-statistic BN renormalises every stage to unit scale, so operand rounding noise is
        carried at full scale through all five stages -- bf16 operands give 42 dB on the stress fixture
        (predicted by CPU emulation with the same rounding points, measured 42.4), fp16 60 dB."""
        value = (self._precision or os.environ.get("V2W_PRECISION", "auto")).lower()
        return "fp16" if value == "auto" else value

    @precision.setter
    def precision(self, value: str) -> None:
        if value not in ("auto", "bf16", "fp16", "fp32", "fp32tc"):
            raise ValueError("precision must be 'auto', 'bf16', 'fp16', 'fp32tc' or 'fp32'")
        self._precision = value

    def engine(self):
        """The libv2w-backed executor of this module (created on first use)."""
        eng = self.__dict__.get("_engine")
        if eng is None:
            from .engine import NativeEngine
            eng = NativeEngine(self)
            object.__setattr__(self, "_engine", eng)
        return eng

    def __getstate__(self):
        # torch.save(module) / copy.deepcopy: the native engine holds ctypes handles and device
        # workspaces; it is rebuilt on first use (only state_dict() is the portable format anyway)
        state = self.__dict__.copy()
        state["_engine"] = None
        return state

    def invalidate(self) -> None:
        """Force a re-fold of the weights on the next native forward (normally automatic:
        parameter versions and storage pointers are checked on every call)."""
        eng = self.__dict__.get("_engine")
        if eng is not None:
            eng.invalidate()

    def _apply(self, fn, *args, **kwargs):
        # .to()/.cuda()/.float(): parameter storage moves and buffer objects are replaced
        out = super()._apply(fn, *args, **kwargs)
        self.invalidate()
        return out

    def conv_modules(self):
        """Weight-normed convs in the order of v2w_params.convs (include/v2w.h)."""
        convs = [self.conv_pre, *self.ups]
        for rb in self.resblocks:
            convs += rb.exec_order()
        convs.append(self.conv_post)
        return convs

    # ------------------------------------------------------------------------------------------
    # forward
    # ------------------------------------------------------------------------------------------
    def forward(self, x, spk_emb=None, noise=None):
        """x [B, num_wv_feat, T] channels-first; spk_emb [B, spk_dim]; noise [B, noise_dim]
        -> [B, 1, 320*T] in (-1, 1)   (vec2wav/models.py:116-147).

        ``noise=None`` draws N(0,1) like the call sites do (vec2wav/train.py:165,259); the
        reference itself would raise there (torch.cat with None)."""
        if spk_emb is None:
            raise TypeError("spk_emb is required (vec2wav/models.py:120 concatenates it)")
        if noise is None:
            noise = torch.randn(x.shape[0], self.h.noise_dim, device=x.device, dtype=x.dtype)
        if x.dim() == 3 and x.shape[0] == 0 and not torch.is_grad_enabled():
            # empty batch: nothing to launch (the reference's convolutions return an empty tensor too)
            return x.new_zeros((0, 1, x.shape[2] * self.engine().total_upsample()), dtype=torch.float32)
        if self.training and not torch.is_grad_enabled() and x.is_cuda and not os.environ.get("V2W_NO_NATIVE_TRAIN"):
            # train() under no_grad (validation / warm-up loops): batch-statistic BN and the power
            # iteration run on the device too
            return self.engine().forward_train(x, spk_emb, noise)
        if self.training or torch.is_grad_enabled():
            if not self.training and not self._warned_grad:
                self._warned_grad = True
                warnings.warn("Generator.eval() called with autograd enabled: running the PyTorch "
                              "autograd delegate. Wrap generation in torch.no_grad() to use the "
                              "native sm_100a path.", RuntimeWarning, stacklevel=2)
            return self._forward_autograd(x, spk_emb, noise)
        return self.engine().forward(x, spk_emb, noise)

    def _forward_autograd(self, x, spk_emb, noise):
        cond = torch.cat((spk_emb, noise), dim=1)
        x = self.conv_pre(x)
        nk = self.num_kernels
        for i, (up, fc, cbn) in enumerate(zip(self.ups, self.fcs, self.cbns)):
            x = up(F.leaky_relu(x, LRELU_SLOPE))
            x = cbn(x, fc(cond))
            acc = self.resblocks[i * nk](x)
            for j in range(1, nk):
                acc += self.resblocks[i * nk + j](x)
            x = acc / nk
        # default slope 0.01 here, not LRELU_SLOPE (vec2wav/models.py:143)
        return torch.tanh(self.conv_post(F.leaky_relu(x)))

    def remove_weight_norm(self):
        """Strip the weight-norm hooks: ``weight_g/weight_v`` -> ``weight`` (models.py:149-156)."""
        print('Removing weight norm...')
        for up in self.ups:
            _strip_wn(up)
        for rb in self.resblocks:
            rb.remove_weight_norm()
        _strip_wn(self.conv_pre)
        _strip_wn(self.conv_post)
        self.invalidate()
