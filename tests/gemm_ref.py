"""Reference arithmetic and launch helper for v2w_tapgemm unit tests (TEST INFRASTRUCTURE)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from wavthruvec_pytorch_b200 import native


def tapgemm_expected(A, W, tiles, n_tile, bias, scale, shift, residual, accum, out_scale, cmod):
    """float64 evaluation of the tap-GEMM contract (include/v2w.h, v2w_layer_desc).
    A [B,rows,cin]; W [taps][n_total][cin]; tiles = list of [(tap_row, offset), ...]."""
    A = np.asarray(A, np.float64)
    W = np.asarray(W, np.float64)
    B, rows, _ = A.shape
    n_total = W.shape[1]
    acc = np.zeros((B, rows, n_total))
    for t, taps in enumerate(tiles):
        cols = slice(t * n_tile, (t + 1) * n_tile)
        for j, off in taps:
            sh = np.zeros_like(A)
            lo, hi = max(0, -off), min(rows, rows - off)
            if hi > lo:
                sh[:, lo:hi] = A[:, lo + off: hi + off]
            acc[:, :, cols] += sh @ W[j][cols].T
    n_idx = np.arange(n_total) % cmod
    v = acc
    if bias is not None:
        v = v + np.asarray(bias, np.float64)[n_idx]
    if scale is not None:
        v = v * np.asarray(scale, np.float64)[:, None, n_idx] + np.asarray(shift, np.float64)[:, None, n_idx]
    if residual is not None:
        v = v + residual
    if accum is not None:
        v = v + accum
    return v * out_scale


def run_tapgemm(engine, A, W, tiles, n_tile, cmod, bias=None, scale=None, shift=None,
                residual=None, accum=None, out_scale=1.0, out_slope=0.1, in_slope=1.0,
                tc_flags=0, device="cuda", want_f32=True, want_bf16=True, in_bf16=True):
    """Launch v2w_tapgemm on `device`.  A/W are fp32 numpy arrays holding the values the
    kernel should see (bf16-representable when the engine consumes bf16).  Returns
    (out_f32 or None, out_bf16-as-fp32 or None)."""
    lib = native.load()
    dev = torch.device(device)
    B, rows, cin = A.shape
    taps_n, n_total, _ = W.shape
    keep = []

    def up(a, dtype=torch.float32):
        if a is None:
            return None
        tt = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev).to(dtype).contiguous()
        keep.append(tt)
        return tt

    a_dev = up(A, torch.bfloat16 if in_bf16 else torch.float32)
    w16 = up(W, torch.bfloat16)
    w32 = up(np.ascontiguousarray(W.transpose(0, 2, 1)))     # [taps][cin][n_total]
    args = native.TapGemmArgs()
    args.engine = engine
    args.batch, args.rows, args.cin, args.n_total, args.cmod = B, rows, cin, n_total, cmod
    args.n_tile, args.n_tiles = n_tile, n_total // n_tile
    for t, taps in enumerate(tiles):
        args.num_taps[t] = len(taps)
        for jj, (j, off) in enumerate(taps):
            args.tap_row[t][jj] = j
            args.tap_off[t][jj] = off
    args.in_act = a_dev.data_ptr()
    args.in_is_bf16 = 1 if in_bf16 else 0
    args.in_slope = in_slope
    args.w16 = w16.data_ptr()
    args.w32 = w32.data_ptr()
    for name, val in (("bias", bias), ("scale", scale), ("shift", shift),
                      ("residual", residual), ("accum_in", accum)):
        tt = up(val)
        setattr(args, name, tt.data_ptr() if tt is not None else None)
    args.out_scale = out_scale
    o32 = torch.full((B, rows, n_total), float("nan"), device=dev) if want_f32 else None
    o16 = torch.full((B, rows, n_total), float("nan"), device=dev, dtype=torch.bfloat16) if want_bf16 else None
    args.out_f32 = o32.data_ptr() if want_f32 else None
    args.out_bf16 = o16.data_ptr() if want_bf16 else None
    args.out_slope = out_slope
    args.tc_flags = tc_flags
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        native.check(lib.v2w_tapgemm(C.byref(args), C.c_void_p(stream)))
        torch.cuda.synchronize(dev)
    return (o32.cpu().numpy() if want_f32 else None,
            o16.float().cpu().numpy() if want_bf16 else None)


def conv_tiles(k, d, n_tiles=1):
    pad = (k * d - d) // 2
    return [[(j, j * d - pad) for j in range(k)] for _ in range(n_tiles)]
