"""The reference's validation step (vec2wav/train.py:246-288) on the native path: glue logic on CPU with
stand-ins for the two library calls, the real thing on the GPU against the reference's op sequence."""
import numpy as np
import pytest
import torch

from oracle import fixtures, mel_numpy
from wavthruvec_pytorch_b200 import validation

from helpers import dropin_in_state, load_golden, t


class _FakeGen:
    """Shape-faithful stand-in: [B, C, T] -> [B, 1, 320 T]."""
    training = False

    class h:
        noise_dim = 4

    def __init__(self):
        self.calls = 0

    def __call__(self, x, spk, noise):
        self.calls += 1
        return torch.tanh(x.mean(dim=1, keepdim=True).repeat_interleave(320, dim=2) + noise.mean())

    def eval(self):
        self.training = False
        return self

    def train(self):
        self.training = True
        return self


def _numpy_mel(y, *args):
    return torch.from_numpy(mel_numpy.mel_spectrogram(y.numpy(), *args).astype(np.float32))


def test_mel_spec_error_glue_matches_the_reference_steps():
    gen = _FakeGen()
    x, s = torch.randn(2, 8, 10), torch.randn(2, 3)
    noise = torch.randn(2, 4)
    frames = 10 * 320 // 256                         # 12 generated frames
    y_mel = torch.randn(2, frames + 3, 80)           # ground truth is longer: truncated like train.py:273
    err, wav = validation.mel_spec_error(gen, x, s, y_mel, noise=noise, mel_fn=_numpy_mel)
    assert tuple(wav.shape) == (2, 1, 3200)
    want_mel = _numpy_mel(wav.squeeze(1), 1024, 80, 16000, 256, 1024, 0, None).permute(0, 2, 1)
    want = torch.nn.functional.l1_loss(y_mel[:, :frames], want_mel).item()
    assert abs(err - want) < 1e-6
    with pytest.raises(RuntimeError, match="does not cover"):
        validation.mel_spec_error(gen, x, s, y_mel[:, :5], noise=noise, mel_fn=_numpy_mel)
    gen.train()
    with pytest.raises(RuntimeError, match="eval"):
        validation.mel_spec_error(gen, x, s, y_mel, noise=noise, mel_fn=_numpy_mel)
    with pytest.raises(ValueError):
        validation.validate(gen, [])
    assert gen.training                               # validate() restores the mode it found


@pytest.mark.gpu
def test_validation_step_on_the_gpu_matches_the_reference_ops():
    torch.backends.cudnn.allow_tf32 = False           # the delegate is the fp32 reference here
    torch.backends.cuda.matmul.allow_tf32 = False
    g = load_golden("rb2_768_stress")
    gen = dropin_in_state(g, cache=False).to("cuda:0").eval()
    gen.precision = "fp32"
    x, s, n = (t(a).cuda() for a in (g.x, g.spk, g.noise))
    with torch.no_grad():
        ref_wav = gen._forward_autograd(x, s, n)
    ref_mel = mel_numpy.mel_spectrogram(ref_wav.squeeze(1).cpu().numpy(), 1024, 80, 16000, 256, 1024, 0, None)
    y_mel = torch.from_numpy(ref_mel.astype(np.float32)).permute(0, 2, 1) + 0.25      # known offset
    err, wav = validation.mel_spec_error(gen, x, s, y_mel.cuda(), noise=n)
    assert np.abs((wav - ref_wav).cpu().numpy()).max() <= 1e-4
    assert abs(err - 0.25) < 2e-3
    gen.train()
    mean = validation.validate(gen, [(x, s, y_mel.cuda())] * 2)
    assert gen.training and 0.0 <= mean < 20.0        # fresh noise per item: only sanity here
