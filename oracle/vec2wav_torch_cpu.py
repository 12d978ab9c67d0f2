"""torch-CPU port of the reference Generator forward (TEST INFRASTRUCTURE / CPU BASELINE ONLY).

The reference's CPU implementation of the hot path *is* stock PyTorch (ATen -> oneDNN/MKL
convolutions: SURVEY.md section 8c "where the arithmetic lives"), and /root/reference does not
exist on the GPU box.  This file restates `Generator.forward` (vec2wav/models.py:116-147, eval
semantics) as a pure function of a state_dict with the SAME ATen calls in the same order --
including the per-forward weight-norm recomputation the reference's forward-pre hooks perform
 -- so that it (a) matches the reference bit-for-bit on the committed goldens and
(b) is a fair stand-in when bench.py times "the reference's CPU path" on the box's host cores.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.1


def _weight(sd, prefix):
    if prefix + ".weight_v" in sd:
        # what WeightNorm.compute_weight does on every forward (torch._weight_norm, dim=0)
        return torch._weight_norm(sd[prefix + ".weight_v"], sd[prefix + ".weight_g"], 0)
    return sd[prefix + ".weight"]


def _pad(k, d):
    return (k * d - d) // 2


def _cbn(x, z, sd, prefix):
    # BatchNorm1d(affine=False).eval(): running stats (vec2wav/modules.py:23)
    xh = F.batch_norm(x, sd[prefix + ".batch_nrom.running_mean"],
                      sd[prefix + ".batch_nrom.running_var"], None, None, False, 0.1, 1e-5)
    W = sd[prefix + ".layer.weight_orig"]
    # SpectralNorm.compute_weight with do_power_iteration=False (eval)
    sigma = torch.dot(sd[prefix + ".layer.weight_u"], torch.mv(W, sd[prefix + ".layer.weight_v"]))
    gb = F.linear(z, W / sigma, sd[prefix + ".layer.bias"])
    gamma, beta = gb.chunk(2, 1)
    C = x.shape[1]
    return gamma.view(-1, C, 1) * xh + beta.view(-1, C, 1)


@torch.no_grad()
def generator_forward(sd: Dict[str, torch.Tensor], h, x, spk, noise):
    cond = torch.cat((spk, noise), dim=1)
    is_rb1 = h.resblock == '1'
    nk = len(h.resblock_kernel_sizes)
    x = F.conv1d(x, _weight(sd, "conv_pre"), sd["conv_pre.bias"], 1, 3)
    for i, (u, k) in enumerate(zip(h.upsample_rates, h.upsample_kernel_sizes)):
        x = F.leaky_relu(x, LRELU_SLOPE)
        x = F.conv_transpose1d(x, _weight(sd, f"ups.{i}"), sd[f"ups.{i}.bias"], u, (k - u) // 2)
        z = F.linear(cond, sd[f"fcs.{i}.weight"], sd[f"fcs.{i}.bias"])
        x = _cbn(x, z, sd, f"cbns.{i}")
        xs = None
        for j, (rk, rd) in enumerate(zip(h.resblock_kernel_sizes, h.resblock_dilation_sizes)):
            p = f"resblocks.{i * nk + j}"
            y = x
            if is_rb1:
                for s in range(3):
                    t = F.conv1d(F.leaky_relu(y, LRELU_SLOPE), _weight(sd, f"{p}.convs1.{s}"),
                                 sd[f"{p}.convs1.{s}.bias"], 1, _pad(rk, rd[s]), rd[s])
                    t = F.conv1d(F.leaky_relu(t, LRELU_SLOPE), _weight(sd, f"{p}.convs2.{s}"),
                                 sd[f"{p}.convs2.{s}.bias"], 1, _pad(rk, 1), 1)
                    y = t + y
            else:
                for s in range(2):
                    t = F.conv1d(F.leaky_relu(y, LRELU_SLOPE), _weight(sd, f"{p}.convs.{s}"),
                                 sd[f"{p}.convs.{s}.bias"], 1, _pad(rk, rd[s]), rd[s])
                    y = t + y
            if xs is None:
                xs = y
            else:
                xs += y
        x = xs / nk
    x = F.leaky_relu(x)
    return torch.tanh(F.conv1d(x, _weight(sd, "conv_post"), sd["conv_post.bias"], 1, 3))
