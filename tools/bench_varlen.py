"""Variable-length batch: valid audio-seconds per second, zero-padded batch (what the reference's
collate does) against length-bucketed ragged generation (wavthruvec_pytorch_b200.varlen).

    python tools/bench_varlen.py [--batch 64] [--min-frames 50] [--max-frames 500] [--out x.json]

Both arms produce the same valid samples (tests/test_varlen.py); timing is CUDA events around the
whole call (bucket planning, torch.cat of the feature rows and all launches included), L2 flushed
between iterations.
"""
from __future__ import annotations

import argparse
import json
import sys
import warnings
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
warnings.filterwarnings("ignore")


def main():
    import numpy as np
    import torch
    from wavthruvec_pytorch_b200 import Generator, HParams, varlen
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--min-frames", type=int, default=50)
    ap.add_argument("--max-frames", type=int, default=500)
    ap.add_argument("--buckets", type=int, default=4)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.manual_seed(1234)
    h = HParams()
    gen = Generator(h).eval().to(dev)
    rng = np.random.default_rng(1234)
    lengths = [int(v) for v in rng.integers(args.min_frames, args.max_frames + 1, args.batch)]
    feats = [torch.randn(n, h.num_wv_feat, device=dev) for n in lengths]
    spk = torch.randn(args.batch, h.spk_dim, device=dev)
    noise = torch.randn(args.batch, h.noise_dim, device=dev)
    t_max = max(lengths)
    x_pad = torch.zeros(args.batch, h.num_wv_feat, t_max, device=dev)
    for b, f in enumerate(feats):
        x_pad[b, :, :f.shape[0]] = f.T
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    valid_s = sum(lengths) / 50.0

    def timed(fn):
        with torch.no_grad():
            for _ in range(args.warmup):
                fn()
            ms = []
            for _ in range(args.steps):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                torch.cuda.synchronize()
                ms.append(a.elapsed_time(b))
        return float(np.median(ms))

    ms_pad = timed(lambda: gen(x_pad, spk, noise))
    ms_skip = timed(lambda: varlen.generate_varlen(gen, feats, spk, noise, max_buckets=1))
    ms_rag = timed(lambda: varlen.generate_varlen(gen, feats, spk, noise, max_buckets=args.buckets))
    ms_pcm = timed(lambda: varlen.generate_varlen(gen, feats, spk, noise, pcm16=True, max_buckets=1))
    buckets = varlen.plan_buckets(lengths, gen.engine().halo_frames(), args.buckets)
    padded, ragged = varlen.padded_frames(lengths, buckets)
    res = {
        "workload": f"{args.batch} utterances, {args.min_frames}-{args.max_frames} frames uniform (seed 1234), "
                    f"768-ch features, ResBlock2, bf16 operands",
        "valid_audio_s": valid_s, "halo_frames": gen.engine().halo_frames(),
        "buckets": [{"utterances": len(idx), "plan_frames": t} for idx, t in buckets],
        "frames_computed": {"padded": padded, "bucketed": ragged, "valid": sum(lengths)},
        "padded": {"ms": ms_pad, "valid_audio_s_per_s": valid_s / ms_pad * 1e3},
        "ragged_f32": {"ms": ms_skip, "valid_audio_s_per_s": valid_s / ms_skip * 1e3},
        "ragged_pcm16": {"ms": ms_pcm, "valid_audio_s_per_s": valid_s / ms_pcm * 1e3},
        "bucketed_f32": {"ms": ms_rag, "valid_audio_s_per_s": valid_s / ms_rag * 1e3, "max_buckets": args.buckets},
    }
    print(json.dumps(res))
    if args.out:
        Path(args.out).parent.mkdir(parents=True, exist_ok=True)
        Path(args.out).write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
