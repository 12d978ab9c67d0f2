"""Conditional batch-norm container (state_dict contract of vec2wav/modules.py:5-30).

The module owns the parameters/buffers under exactly the reference's names -- including the
misspelt ``batch_nrom`` -- so checkpoints load unchanged:

    cbns.i.batch_nrom.{running_mean,running_var,num_batches_tracked}
    cbns.i.layer.{bias,weight_orig,weight_u,weight_v}

In generation (``eval()``) the arithmetic does not run here: BN running statistics, the
spectral-norm sigma and the gamma/beta projection collapse into one per-(utterance, channel)
affine that the native library folds into the epilogue of the preceding transposed conv
("cond affine").  ``forward`` below is the autograd/train-mode delegate only.
"""
from __future__ import annotations

import torch
import torch.nn as nn
from torch.nn.utils import spectral_norm

Z_CHANNELS = 128  # hard-coded in the reference (vec2wav/modules.py:9, vec2wav/models.py:110)


class ConditionalBatchNorm1d(nn.Module):
    def __init__(self, num_features: int, z_channels: int = Z_CHANNELS):
        super().__init__()
        self.num_features = num_features
        self.z_channels = z_channels
        # attribute name is part of the checkpoint format (vec2wav/modules.py:14)
        self.batch_nrom = nn.BatchNorm1d(num_features, affine=False)
        proj = nn.Linear(z_channels, 2 * num_features)
        self.layer = spectral_norm(proj)
        # vec2wav/modules.py:17-18: after spectral_norm, ``layer.weight`` aliases the storage
        # of ``weight_orig``, so this draw *is* the scale init N(1, 0.02); bias starts at 0.
        self.layer.weight.data.normal_(1.0, 0.02)
        self.layer.bias.data.zero_()

    def forward(self, inputs: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
        """Train-mode / autograd delegate: BN (batch stats in train) then gamma*x+beta.

        gamma is the first ``num_features`` outputs of the projection (vec2wav/modules.py:24).
        """
        normed = self.batch_nrom(inputs)
        gb = self.layer(z)
        gamma = gb[:, : self.num_features].unsqueeze(-1)
        beta = gb[:, self.num_features:].unsqueeze(-1)
        return gamma * normed + beta
