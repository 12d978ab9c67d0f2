"""Log-mel consumer: v2w_mel_forward against the reference's op sequence
(torch.stft etc., vec2wav/dataset.py:53-77) on the same GPU.  Prints one JSON line."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from wavthruvec_pytorch_b200 import mel  # noqa: E402

P = dict(n_fft=1024, num_mels=80, sampling_rate=16000, hop_size=256, win_size=1024, fmin=0, fmax=None)


def timed(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def torch_mel(y, fb, win):
    pad = (P["n_fft"] - P["hop_size"]) // 2
    yp = torch.nn.functional.pad(y.unsqueeze(1), (pad, pad), mode="reflect").squeeze(1)
    spec = torch.stft(yp, P["n_fft"], hop_length=P["hop_size"], win_length=P["win_size"], window=win,
                      center=False, normalized=False, onesided=True, return_complex=True)
    mag = torch.sqrt(spec.real ** 2 + spec.imag ** 2 + 1e-9)
    return torch.log(torch.clamp(torch.matmul(fb, mag), min=1e-5))


def main():
    out = {}
    fb = torch.from_numpy(mel.filterbank(1024, 80, 16000, 0, None)).cuda()
    win = torch.hann_window(1024, device="cuda")
    for name, (B, L) in {"cfg2": (16, 64000), "cfg5_per_gpu": (64, 160000)}.items():
        y = (torch.rand(B, L, device="cuda") * 2 - 1) * 0.5
        ms = timed(lambda: mel.mel_spectrogram(y, **P))
        ms_t = timed(lambda: torch_mel(y, fb, win))
        err = (mel.mel_spectrogram(y, **P) - torch_mel(y, fb, win)).abs().max().item()
        rows = L // 256 + 3
        flop = 2.0 * B * rows * 1152 * 3072
        out[name] = {"native_ms": ms, "torch_stft_ms": ms_t, "audio_s_per_s": B * L / 16000 / (ms * 1e-3),
                     "stft_gemm_tflops_if_all_time": flop / (ms * 1e-3) / 1e12, "max_abs_vs_torch": err}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
