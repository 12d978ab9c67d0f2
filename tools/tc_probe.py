"""Hardware probe of the tcgen05 tap-GEMM (run on the GPU box).

Runs a grid of (shape, tap pattern, descriptor variant) cases of v2w_tapgemm, each in its own
subprocess so that a trapped kernel (sticky context error) cannot poison the next case, and
writes one JSON line per case to gpurun_out/tc_probe.jsonl.

    python tools/tc_probe.py            # the whole grid
    python tools/tc_probe.py --one '<json case>'
"""
from __future__ import annotations

import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def one(case):
    import numpy as np
    import gemm_ref
    import plan_interp
    from wavthruvec_pytorch_b200 import native
    rng = np.random.default_rng(case.get("seed", 0))
    B, rows, cin, n_total, n_tile = case["B"], case["rows"], case["cin"], case["n_total"], case["n_tile"]
    k, d = case["k"], case["d"]
    A = plan_interp.bf16_round(rng.standard_normal((B, rows, cin)).astype(np.float32))
    W = plan_interp.bf16_round((rng.standard_normal((k, n_total, cin)) / np.sqrt(cin * k)).astype(np.float32))
    tiles = gemm_ref.conv_tiles(k, d, n_total // n_tile)
    cmod = case.get("cmod", n_total)
    full = case.get("epilogue", False)
    bias = rng.standard_normal(cmod).astype(np.float32) if full else None
    scale = rng.standard_normal((B, cmod)).astype(np.float32) if full else None
    shift = rng.standard_normal((B, cmod)).astype(np.float32) if full else None
    resid = rng.standard_normal((B, rows, n_total)).astype(np.float32) if full else None
    accum = rng.standard_normal((B, rows, n_total)).astype(np.float32) if full else None
    out_scale = 1.0 / 3.0 if full else 1.0
    want = gemm_ref.tapgemm_expected(A, W, tiles, n_tile, bias, scale, shift, resid, accum, out_scale, cmod)
    import os
    if case.get("msub"):
        os.environ["V2W_TC_MSUB"] = str(case["msub"])
    else:
        os.environ.pop("V2W_TC_MSUB", None)
    engine = native.ENGINE_TC if case.get("engine", "tc") == "tc" else native.ENGINE_SIMT
    o32, o16 = gemm_ref.run_tapgemm(engine, A, W, tiles, n_tile, cmod, bias, scale, shift, resid,
                                    accum, out_scale, 0.1, tc_flags=case.get("flags", 0))
    err = float(np.abs(o32 - want).max())
    ref = float(np.abs(want).max())
    lre = np.where(want > 0, want, want * 0.1)
    err16 = float(np.abs(o16 - lre).max())
    bad_rows = np.unique(np.argwhere(np.abs(o32 - want) > 1e-3 * max(1.0, ref))[:, 1])[:12].tolist()
    return dict(case=case, max_err=err, ref_max=ref, max_err_bf16=err16,
                nan=int(np.isnan(o32).sum()), bad_rows=bad_rows,
                ok=bool(err < 2e-3 * max(1.0, ref) and err16 < 2e-2 * max(1.0, ref)))


def grid():
    cases = []
    base = dict(B=2, rows=300, cin=64, n_total=64, n_tile=64, k=3, d=1)
    # super-tile heights on the simplest halo case, then without halo (k=1)
    for msub in (1, 2, 4):
        cases.append(dict(base, rows=1100, msub=msub, tag="msub"))
        cases.append(dict(base, rows=1100, cin=16, n_total=16, n_tile=16, k=11, d=5, msub=msub,
                          epilogue=True, tag="msub"))
    cases.append(dict(base, k=1, flags=0, tag="k1"))
    for cin, n in ((16, 16), (32, 32), (64, 64), (128, 128), (256, 256), (768, 512)):
        for (k, d) in ((3, 1), (7, 3), (11, 3), (11, 5)):
            n_tile = min(n, 256)
            cases.append(dict(B=2, rows=517, cin=cin, n_total=n, n_tile=n_tile, k=k, d=d, flags=0,
                              epilogue=True, tag="shape"))
    for rows in (1, 31, 128, 129, 1000):
        cases.append(dict(B=3, rows=rows, cin=64, n_total=128, n_tile=128, k=7, d=1, flags=0,
                          epilogue=True, tag="rows"))
    cases.append(dict(B=1, rows=40000, cin=32, n_total=32, n_tile=32, k=11, d=3, flags=0,
                      epilogue=True, tag="long"))
    cases.append(dict(B=2, rows=300, cin=64, n_total=64, n_tile=64, k=3, d=1, engine="simt", tag="simt"))
    cases.append(dict(B=2, rows=300, cin=256, n_total=256, n_tile=256, k=11, d=5, engine="simt",
                      epilogue=True, tag="simt"))
    return cases


def worker(start):
    """Run cases[start:] in this process; stop at the first case that raises (a trapped kernel
    leaves a sticky context error) so that the parent can restart after it."""
    cases = grid()
    for i in range(start, len(cases)):
        try:
            res = one(cases[i])
        except Exception as e:  # noqa: BLE001
            res = dict(case=cases[i], ok=False, error=f"{type(e).__name__}: {e}"[:500])
            print("RESULT " + json.dumps(dict(index=i, **res)), flush=True)
            sys.exit(3)
        print("RESULT " + json.dumps(dict(index=i, **res)), flush=True)


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "--from":
        worker(int(sys.argv[2]))
        return
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    cases = grid()
    n_ok, nxt = 0, 0
    with open(out / "tc_probe.jsonl", "w") as f:
        while nxt < len(cases):
            try:
                p = subprocess.run([sys.executable, __file__, "--from", str(nxt)],
                                   capture_output=True, text=True, timeout=600)
                lines = [json.loads(l[7:]) for l in p.stdout.splitlines() if l.startswith("RESULT ")]
                tail = p.stderr[-600:]
            except subprocess.TimeoutExpired as e:
                txt = (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
                lines = [json.loads(l[7:]) for l in txt.splitlines() if l.startswith("RESULT ")]
                tail = "timeout"
            if not lines:
                lines = [dict(index=nxt, case=cases[nxt], ok=False, error="worker died: " + tail)]
            for res in lines:
                f.write(json.dumps(res) + "\n")
                f.flush()
                n_ok += bool(res.get("ok"))
                c = res["case"]
                print(f"{'OK ' if res.get('ok') else 'BAD'} {c.get('tag')} cin={c['cin']} n={c['n_total']} "
                      f"k={c['k']} d={c['d']} rows={c['rows']} msub={c.get('msub', 0)} "
                      f"eng={c.get('engine', 'tc')} err={res.get('max_err')} err16={res.get('max_err_bf16')} "
                      f"nan={res.get('nan')} bad_rows={res.get('bad_rows')} {res.get('error', '')}",
                      flush=True)
            nxt = lines[-1]["index"] + 1
    print(f"{n_ok}/{len(cases)} cases ok")


if __name__ == "__main__":
    main()
