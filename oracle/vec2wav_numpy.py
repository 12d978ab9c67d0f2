"""numpy restatement of Generator.forward, eval() and train()-under-no_grad modes (TEST INFRASTRUCTURE ONLY).

Plain array arithmetic from the formulas -- no torch convolution is involved, so it checks the
*semantics* (padding, transposed-conv indexing, weight-norm axis, gamma/beta order, the two
leaky-relu slopes) independently of ATen.  Defaults to float64.

Each function cites the reference lines it restates.  Input is the module's state_dict as
{name: ndarray}; both weight-normed (`weight_g`/`weight_v`) and stripped (`weight`) forms are
accepted (vec2wav/models.py:149-156).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np

LRELU_SLOPE = 0.1          # vec2wav/models.py:10
LRELU_SLOPE_POST = 0.01    # F.leaky_relu default, vec2wav/models.py:143
BN_EPS = 1e-5              # nn.BatchNorm1d default, vec2wav/modules.py:14


def leaky_relu(x, slope):
    return np.where(x > 0, x, x * slope)


def effective_weight(sd: Dict[str, np.ndarray], prefix: str, dtype) -> np.ndarray:
    """weight_norm(dim=0): w = g * v / ||v||, norm over every dim but 0 -- Cout for Conv1d
    [Cout,Cin,K], but Cin for ConvTranspose1d [Cin,Cout,K] (torch.nn.utils.weight_norm as
    applied at vec2wav/models.py:83,90-92,18-33,58-61,100)."""
    if prefix + ".weight_v" in sd:
        v = sd[prefix + ".weight_v"].astype(dtype)
        g = sd[prefix + ".weight_g"].astype(dtype)
        norm = np.sqrt((v.reshape(v.shape[0], -1) ** 2).sum(axis=1)).reshape(-1, 1, 1)
        return v * (g / norm)
    return sd[prefix + ".weight"].astype(dtype)


def conv1d(x, w, b, dilation: int, padding: int):
    """nn.Conv1d, stride 1.  x [B,Cin,L], w [Cout,Cin,K] -> [B,Cout,L + 2p - d(K-1)]."""
    B, _, L = x.shape
    Cout, _, K = w.shape
    xp = np.pad(x, ((0, 0), (0, 0), (padding, padding)))
    Lout = L + 2 * padding - dilation * (K - 1)
    y = np.zeros((B, Cout, Lout), dtype=x.dtype)
    for k in range(K):
        y += np.einsum("oc,bcl->bol", w[:, :, k], xp[:, :, k * dilation: k * dilation + Lout])
    return y + b.reshape(1, -1, 1)


def conv_transpose1d(x, w, b, stride: int, padding: int):
    """nn.ConvTranspose1d, output_padding 0.  x [B,Cin,L], w [Cin,Cout,K]:
    y[b, co, t*stride + k - padding] += x[b, ci, t] * w[ci, co, k]."""
    B, _, L = x.shape
    _, Cout, K = w.shape
    full = np.zeros((B, Cout, (L - 1) * stride + K), dtype=x.dtype)
    for k in range(K):
        full[:, :, k: k + (L - 1) * stride + 1: stride] += np.einsum("co,bcl->bol", w[:, :, k], x)
    Lout = (L - 1) * stride - 2 * padding + K
    return full[:, :, padding: padding + Lout] + b.reshape(1, -1, 1)


def get_padding(k: int, d: int) -> int:
    return (k * d - d) // 2   # vec2wav/utils.py:35-36


def cbn_eval(x, z, sd, prefix: str, dtype):
    """ConditionalBatchNorm1d.forward in eval() (vec2wav/modules.py:20-30): BN with running
    stats, then gamma*x_hat + beta where [gamma|beta] = SN-Linear(z) and, in eval, the spectral
    norm divides by sigma = u . (W_orig v) with the STORED u, v (no power iteration)."""
    mean = sd[prefix + ".batch_nrom.running_mean"].astype(dtype)
    var = sd[prefix + ".batch_nrom.running_var"].astype(dtype)
    xh = (x - mean.reshape(1, -1, 1)) / np.sqrt(var.reshape(1, -1, 1) + BN_EPS)
    W = sd[prefix + ".layer.weight_orig"].astype(dtype)
    u = sd[prefix + ".layer.weight_u"].astype(dtype)
    v = sd[prefix + ".layer.weight_v"].astype(dtype)
    sigma = u @ (W @ v)
    gb = z @ (W / sigma).T + sd[prefix + ".layer.bias"].astype(dtype)
    C = x.shape[1]
    gamma, beta = gb[:, :C], gb[:, C:]          # chunk(2, 1): gamma first (modules.py:24)
    return gamma[:, :, None] * xh + beta[:, :, None]


def cbn_train(x, z, sd, prefix: str, dtype, momentum: float = 0.1, sn_eps: float = 1e-12):
    """ConditionalBatchNorm1d.forward with module.training set (vec2wav/modules.py:14,16,20-30).
    nn.BatchNorm1d normalises with THIS batch's statistics (biased variance over batch x time) and
    moves the running statistics by `momentum` (unbiased variance); torch.nn.utils.spectral_norm
    runs one power iteration first -- v <- normalize(W^T u), u <- normalize(W v), both stored --
    and divides by sigma = u . (W v).  MUTATES `sd` like the module mutates its buffers."""
    n = x.shape[0] * x.shape[2]
    mean = x.mean(axis=(0, 2))
    var = x.var(axis=(0, 2))                     # biased
    xh = (x - mean.reshape(1, -1, 1)) / np.sqrt(var.reshape(1, -1, 1) + BN_EPS)
    rm, rv = sd[prefix + ".batch_nrom.running_mean"], sd[prefix + ".batch_nrom.running_var"]
    sd[prefix + ".batch_nrom.running_mean"] = ((1 - momentum) * rm.astype(dtype) + momentum * mean).astype(rm.dtype)
    sd[prefix + ".batch_nrom.running_var"] = ((1 - momentum) * rv.astype(dtype)
                                              + momentum * var * n / (n - 1)).astype(rv.dtype)
    sd[prefix + ".batch_nrom.num_batches_tracked"] = sd[prefix + ".batch_nrom.num_batches_tracked"] + 1
    W = sd[prefix + ".layer.weight_orig"].astype(dtype)
    u = sd[prefix + ".layer.weight_u"].astype(dtype)
    v = W.T @ u
    v = v / max(np.sqrt((v * v).sum()), sn_eps)
    u = W @ v
    u = u / max(np.sqrt((u * u).sum()), sn_eps)
    sd[prefix + ".layer.weight_u"] = u.astype(sd[prefix + ".layer.weight_u"].dtype)
    sd[prefix + ".layer.weight_v"] = v.astype(sd[prefix + ".layer.weight_v"].dtype)
    sigma = u @ (W @ v)
    gb = z @ (W / sigma).T + sd[prefix + ".layer.bias"].astype(dtype)
    C = x.shape[1]
    gamma, beta = gb[:, :C], gb[:, C:]
    return gamma[:, :, None] * xh + beta[:, :, None]


def resblock2(x, sd, prefix, k, dil, dtype):
    """ResBlock2.forward (vec2wav/models.py:65-70): dilations dil[0], dil[1] only."""
    for s in range(2):
        w = effective_weight(sd, f"{prefix}.convs.{s}", dtype)
        b = sd[f"{prefix}.convs.{s}.bias"].astype(dtype)
        x = conv1d(leaky_relu(x, LRELU_SLOPE), w, b, dil[s], get_padding(k, dil[s])) + x
    return x


def resblock1(x, sd, prefix, k, dil, dtype):
    """ResBlock1.forward (vec2wav/models.py:37-44)."""
    for s in range(3):
        w1 = effective_weight(sd, f"{prefix}.convs1.{s}", dtype)
        b1 = sd[f"{prefix}.convs1.{s}.bias"].astype(dtype)
        w2 = effective_weight(sd, f"{prefix}.convs2.{s}", dtype)
        b2 = sd[f"{prefix}.convs2.{s}.bias"].astype(dtype)
        y = conv1d(leaky_relu(x, LRELU_SLOPE), w1, b1, dil[s], get_padding(k, dil[s]))
        y = conv1d(leaky_relu(y, LRELU_SLOPE), w2, b2, 1, get_padding(k, 1))
        x = y + x
    return x


def generator_forward(sd: Dict[str, np.ndarray], h, x, spk, noise, dtype=np.float64,
                      return_pre_tanh: bool = False, taps: Optional[List] = None, train: bool = False):
    """Generator.forward (vec2wav/models.py:116-147), eval() semantics; `train=True`: train()
    semantics under no_grad (batch-statistic BN, power iteration), `sd` is updated in place.

    x [B, num_wv_feat, T], spk [B, spk_dim], noise [B, noise_dim] -> [B, 1, 320 T].
    `taps`, if a list, receives (name, ndarray) of intermediate activations."""
    x = np.asarray(x, dtype=dtype)
    cond = np.concatenate([np.asarray(spk, dtype), np.asarray(noise, dtype)], axis=1)  # :120
    is_rb1 = h.resblock == '1'                                                          # :84
    nk = len(h.resblock_kernel_sizes)
    x = conv1d(x, effective_weight(sd, "conv_pre", dtype), sd["conv_pre.bias"].astype(dtype), 1, 3)
    if taps is not None:
        taps.append(("conv_pre", x))
    for i, (u, k) in enumerate(zip(h.upsample_rates, h.upsample_kernel_sizes)):
        x = leaky_relu(x, LRELU_SLOPE)                                                  # :128
        x = conv_transpose1d(x, effective_weight(sd, f"ups.{i}", dtype),
                             sd[f"ups.{i}.bias"].astype(dtype), u, (k - u) // 2)        # :129
        z = cond @ sd[f"fcs.{i}.weight"].astype(dtype).T + sd[f"fcs.{i}.bias"].astype(dtype)
        x = (cbn_train if train else cbn_eval)(x, z, sd, f"cbns.{i}", dtype)            # :134
        if taps is not None:
            taps.append((f"cbn.{i}", x))
        xs = None
        for j, (rk, rd) in enumerate(zip(h.resblock_kernel_sizes, h.resblock_dilation_sizes)):
            prefix = f"resblocks.{i * nk + j}"
            y = (resblock1 if is_rb1 else resblock2)(x, sd, prefix, rk, rd, dtype)
            xs = y if xs is None else xs + y
        x = xs / nk                                                                     # :141
        if taps is not None:
            taps.append((f"stage.{i}", x))
    x = leaky_relu(x, LRELU_SLOPE_POST)                                                 # :143
    pre = conv1d(x, effective_weight(sd, "conv_post", dtype), sd["conv_post.bias"].astype(dtype), 1, 3)
    y = np.tanh(pre)
    return (y, pre) if return_pre_tanh else y


def state_dict_to_numpy(module) -> Dict[str, np.ndarray]:
    return {k: v.detach().cpu().numpy() for k, v in module.state_dict().items()}
