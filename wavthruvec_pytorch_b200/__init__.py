"""B200-native Vec2Wav generator forward pass (drop-in for WavThruVec_pytorch's
vec2wav.models.Generator).  See DESIGN.md / INTEGRATION.md at the repository root."""
from .hparams import FRAMES_PER_SECOND, SAMPLES_PER_FRAME, HParams
from .models import Generator, ResBlock1, ResBlock2
from .modules import ConditionalBatchNorm1d
from .mel import mel_spectrogram

__all__ = ["Generator", "ResBlock1", "ResBlock2", "ConditionalBatchNorm1d", "HParams",
           "FRAMES_PER_SECOND", "SAMPLES_PER_FRAME", "mel_spectrogram"]
