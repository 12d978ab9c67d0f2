"""One train()-mode forward (cfg2) and one log-mel call between cudaProfilerStart/Stop, for
`ncu --profile-from-start off` (tools/gpu_final2.sh)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from wavthruvec_pytorch_b200 import Generator, HParams, mel_spectrogram  # noqa: E402

P = dict(n_fft=1024, num_mels=80, sampling_rate=16000, hop_size=256, win_size=1024, fmin=0, fmax=None)
h = HParams()
torch.manual_seed(1234)
gen = Generator(h).to("cuda:0").train()
x, s, n = (torch.randn(16, h.num_wv_feat, 200, device="cuda"), torch.randn(16, h.spk_dim, device="cuda"),
           torch.randn(16, h.noise_dim, device="cuda"))
with torch.no_grad():
    for _ in range(2):
        y = gen(x, s, n)
        m = mel_spectrogram(y.squeeze(1), **P)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    y = gen(x, s, n)
    m = mel_spectrogram(y.squeeze(1), **P)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("ok", tuple(y.shape), tuple(m.shape))
