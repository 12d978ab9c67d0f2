"""Feature files -> WAV files, end to end: writes a synthetic tree in the
reference's on-disk formats under a scratch directory, then times files.synthesize_files.

    python tools/bench_files.py [--utts 256] [--min-frames 50] [--max-frames 500] [--out x.json]
"""
from __future__ import annotations

import argparse
import json
import shutil
import sys
import tempfile
import time
import warnings
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
warnings.filterwarnings("ignore")


def main():
    import numpy as np
    import torch
    from wavthruvec_pytorch_b200 import Generator, HParams, files
    ap = argparse.ArgumentParser()
    ap.add_argument("--utts", type=int, default=256)
    ap.add_argument("--speakers", type=int, default=16)
    ap.add_argument("--min-frames", type=int, default=50)
    ap.add_argument("--max-frames", type=int, default=500)
    ap.add_argument("--batch-frames", type=int, default=16000)
    ap.add_argument("--readers", type=int, default=4)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.manual_seed(1234)
    h = HParams()
    gen = Generator(h).eval().to(dev)
    rng = np.random.default_rng(1234)
    root = Path(tempfile.mkdtemp(prefix="v2w_files_"))
    try:
        feat_dir, spk_dir, out_dir = root / "w2v_feat", root / "spk_emb", root / "wav"
        spk_dir.mkdir()
        items = []
        t0 = time.perf_counter()
        for s in range(args.speakers):
            (feat_dir / f"SPK{s:04d}").mkdir(parents=True)
            torch.save(torch.randn(1, 1, h.spk_dim), spk_dir / f"SPK{s:04d}.pth")
        total_frames = 0
        for i in range(args.utts):
            spk = f"SPK{i % args.speakers:04d}"
            n = int(rng.integers(args.min_frames, args.max_frames + 1))
            np.save(feat_dir / spk / f"utt{i:05d}.npy", rng.standard_normal((1, n, h.num_wv_feat)).astype(np.float32))
            items.append(files.Item(f"{spk}/utt{i:05d}", spk))
            total_frames += n
        t_write = time.perf_counter() - t0
        # warm-up pass over everything (builds the plans of every batch shape, folds the weights, warms the
        # page cache), then the timed pass: steady-state throughput of a long-running job
        files.synthesize_files(gen, items, str(feat_dir), str(spk_dir), str(root / "warm"),
                               max_batch_frames=args.batch_frames, readers=args.readers)
        st = files.synthesize_files(gen, items, str(feat_dir), str(spk_dir), str(out_dir),
                                    max_batch_frames=args.batch_frames, readers=args.readers)
        wav_bytes = sum(p.stat().st_size for p in out_dir.rglob("*.wav"))
        res = {
            "workload": f"{args.utts} feature files ({args.min_frames}-{args.max_frames} frames, 768-ch, fp32 .npy), "
                        f"{args.speakers} speaker .pth files -> 16-bit WAV files; ResBlock2, bf16 operands",
            "utterances": st.utterances, "batches": st.batches, "audio_s": st.audio_s, "wall_s": st.wall_s,
            "audio_s_per_s": st.audio_s_per_s, "feature_bytes": total_frames * h.num_wv_feat * 4,
            "wav_bytes": wav_bytes, "readers": args.readers, "batch_frames": args.batch_frames,
            "feature_read_GBps": total_frames * h.num_wv_feat * 4 / st.wall_s / 1e9,
            "fixture_write_s": t_write,
        }
        print(json.dumps(res))
        if args.out:
            Path(args.out).parent.mkdir(parents=True, exist_ok=True)
            Path(args.out).write_text(json.dumps(res, indent=1))
    finally:
        shutil.rmtree(root, ignore_errors=True)


if __name__ == "__main__":
    main()
