"""Per-op device times of one plan (CUDA events recorded by libv2w around every launch).

    python tools/op_table.py --config cfg2 [--resblock 1] [--precision bf16] [--out gpurun_out/x.json]

Prints one row per op: layer geometry, tile shape, ms, achieved TFLOP/s and the algorithmic
HBM bytes/ms of the op's activations (reads + writes as the kernel performs them).
"""
from __future__ import annotations

import argparse
import json
import sys
import warnings
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
warnings.filterwarnings("ignore")


def op_bytes(op, L, prec_bf16):
    """Bytes the op's kernel moves through global memory for activations (no weights)."""
    elems_in = op.batch * op.rows * L.cin
    elems_out = op.batch * op.rows * L.n_total
    b = elems_in * (2 if op.in_is_bf16 else 4)
    for off, sz in ((op.residual, 4), (op.accum_in, 4), (op.out_f32, 4), (op.out_bf16, 2)):
        if off >= 0:
            b += elems_out * sz
    return b


def main():
    import torch
    from bench import CONFIGS, make_hparams
    from wavthruvec_pytorch_b200 import Generator, native
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--resblock", default="2")
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    batch, frames, feat = CONFIGS[args.config]
    h = make_hparams(args, feat)
    torch.manual_seed(1234)
    dev = torch.device("cuda:0")
    gen = Generator(h).eval().to(dev)
    gen.precision = args.precision
    eng = gen.engine()
    x = torch.randn(batch, feat, frames, device=dev)
    s = torch.randn(batch, 192, device=dev)
    n = torch.randn(batch, 192, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    with torch.no_grad():
        for _ in range(3):
            gen(x, s, n)
        prof = eng.profile_ops(x, s, n, reps=args.reps, before_each=lambda: flush.zero_())
    layers = eng.layers()
    rows = []
    tot = 0.0
    for i, (op, ms) in enumerate(prof):
        tot += ms
        if op.kind == native.OP_TAPGEMM:
            L = layers[op.layer]
            by = op_bytes(op, L, args.precision == "bf16")
            rows.append(dict(i=i, kind="gemm", role=L.role, stage=L.stage, cin=L.cin, n=L.n_total,
                             k=L.kernel_size, d=L.dilation, up=L.up, rows=op.rows, n_tile=op.n_tile,
                             taps=[op.num_taps[t] for t in range(op.n_tiles)], ms=ms,
                             tflops=op.flops / ms / 1e9, gbs=by / ms / 1e6, mb=by / 1e6,
                             resid=op.residual >= 0, accum=op.accum_in >= 0,
                             f32=op.out_f32 >= 0, b16=op.out_bf16 >= 0))
        elif op.kind in (native.OP_MRF_FUSED, native.OP_PAIR_FUSED):
            C = op.n_tile
            by = op.batch * op.rows * C * (2 + 4 + (4 if op.out_f32 >= 0 else 0) + (2 if op.out_bf16 >= 0 else 0))
            rows.append(dict(i=i, kind="fused", stage=op.stage, cin=C, rows=op.rows, ms=ms,
                             tflops=op.flops / ms / 1e9, gbs=by / ms / 1e6, mb=by / 1e6))
        elif op.kind in (native.OP_MRF_WIDE1, native.OP_MRF_WIDE2):
            C = op.n_tile
            half = 1 if op.kind == native.OP_MRF_WIDE1 else 2
            by = op.batch * op.rows * C * ((2 + 4 + 3 * 2) if half == 1 else (4 * 2 + 4 + 2))
            rows.append(dict(i=i, kind="fused", stage=op.stage, cin=C, rows=op.rows, ms=ms, wide=half,
                             tflops=op.flops / ms / 1e9, gbs=by / ms / 1e6, mb=by / 1e6))
        else:
            rows.append(dict(i=i, kind={0: "transpose", 1: "cond", 3: "post"}[op.kind], ms=ms))
    for r in rows:
        if r["kind"] == "gemm":
            print(f"{r['i']:3d} gemm role={r['role']} st={r['stage']:2d} cin={r['cin']:4d} n={r['n']:4d} "
                  f"k={r['k']:2d} d={r['d']} up={r['up']} rows={r['rows']:6d} tile={r['n_tile']:3d} "
                  f"taps={r['taps']} {r['ms'] * 1e3:8.1f} us {r['tflops']:7.1f} TF/s "
                  f"{r['gbs']:7.0f} GB/s ({r['mb']:.0f} MB)")
        elif r["kind"] == "fused":
            tag = f"WIDE{r['wide']}" if r.get("wide") else "FUSED"
            print(f"{r['i']:3d} {tag} MRF stage {r['stage']} C={r['cin']:3d} rows={r['rows']:6d} "
                  f"{r['ms'] * 1e3:8.1f} us {r['tflops']:7.1f} TF/s {r['gbs']:7.0f} GB/s ({r['mb']:.0f} MB)")
        else:
            print(f"{r['i']:3d} {r['kind']:9s} {r['ms'] * 1e3:8.1f} us")
    print(f"total {tot:.3f} ms for {batch * frames / 50:.0f} audio-s -> {batch * frames / 50 / tot * 1e3:.0f} audio-s/s")
    if args.out:
        Path(args.out).parent.mkdir(parents=True, exist_ok=True)
        Path(args.out).write_text(json.dumps(dict(config=args.config, resblock=args.resblock,
                                                  precision=args.precision, total_ms=tot, ops=rows)))


if __name__ == "__main__":
    main()
