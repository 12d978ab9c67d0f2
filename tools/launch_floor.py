"""Fixed cost of one generic tap-GEMM launch and the marginal cost of a tile (CUDA events around single
v2w_tapgemm launches).  usage: python tools/launch_floor.py"""
import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from wavthruvec_pytorch_b200 import native  # noqa: E402


def case(B, rows, cin, n_total, n_tile, k, d, resid=False, f32=True, b16=True, reps=12, affine=False):
    lib = native.load()
    dev = torch.device("cuda:0")
    A = torch.randn(B, rows, cin, device=dev).to(torch.bfloat16)
    W = (torch.randn(k, n_total, cin, device=dev) / (cin * k) ** 0.5).to(torch.bfloat16)
    bias = torch.randn(n_total, device=dev)
    res = torch.randn(B, rows, n_total, device=dev) if resid else None
    o32 = torch.empty(B, rows, n_total, device=dev) if f32 else None
    o16 = torch.empty(B, rows, n_total, device=dev, dtype=torch.bfloat16) if b16 else None
    a = native.TapGemmArgs()
    a.engine = native.ENGINE_TC
    a.batch, a.rows, a.cin, a.n_total, a.cmod = B, rows, cin, n_total, n_total
    a.n_tile, a.n_tiles = n_tile, n_total // n_tile
    pad = (k * d - d) // 2
    for t in range(a.n_tiles):
        a.num_taps[t] = k
        for j in range(k):
            a.tap_row[t][j] = j
            a.tap_off[t][j] = j * d - pad
    a.in_act, a.in_is_bf16, a.in_slope = A.data_ptr(), 1, 1.0
    a.w16 = W.data_ptr()
    a.bias = bias.data_ptr()
    if affine:   # per-(utterance, channel) scale / shift: the conditional batch norm of a transposed conv's epilogue
        scale = torch.randn(B, n_total, device=dev)
        shift = torch.randn(B, n_total, device=dev)
        a.scale, a.shift = scale.data_ptr(), shift.data_ptr()
    a.residual = res.data_ptr() if resid else None
    a.out_scale, a.out_slope = 1.0, 0.1
    a.out_f32 = o32.data_ptr() if f32 else None
    a.out_bf16 = o16.data_ptr() if b16 else None
    st = torch.cuda.current_stream().cuda_stream
    times = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        native.check(lib.v2w_tapgemm(C.byref(a), C.c_void_p(st)))
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) * 1e3)
    times = sorted(times[2:])
    flops = 2.0 * B * rows * cin * n_total * k
    med = times[len(times) // 2]
    return dict(B=B, rows=rows, cin=cin, n=n_total, n_tile=n_tile, k=k, us_med=round(med, 2), us_min=round(times[0], 2),
                tflops=round(flops / med / 1e6, 1), resid=resid, f32=f32, b16=b16)


def main():
    out = []
    for args in [
        (1, 128, 64, 64, 64, 1, 1),              # one single-CTA tile, 4 MMAs
        (1, 256, 64, 128, 128, 1, 1),            # one CTA-pair tile
        (74, 256, 64, 128, 128, 1, 1),           # one tile per pair
        (74, 256, 256, 256, 256, 3, 1),          # one tile per pair, 48 MMA steps (stage-0 conv, k = 3)
        (74, 512, 256, 256, 256, 3, 1),          # two tiles per pair
        (74, 1024, 256, 256, 256, 3, 1),         # four tiles per pair
        (74, 2048, 256, 256, 256, 3, 1),         # eight
        (74, 256, 256, 256, 256, 11, 1),         # one tile per pair, 176 MMA steps
        (74, 1024, 256, 256, 256, 11, 1),
        (16, 200, 768, 512, 128, 7, 1),          # conv_pre at cfg2
    ]:
        for resid, f32, b16 in ((False, False, True), (True, True, True)):
            r = case(*args, resid=resid, f32=f32, b16=b16)
            out.append(r)
            print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
