"""train()-mode forward under no_grad: native v2w_forward_train against the
PyTorch delegate on the same GPU, cfg2 shapes.  Prints one JSON line."""
import json
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from wavthruvec_pytorch_b200 import Generator, HParams  # noqa: E402


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    B, T = 16, 200
    h = HParams()
    torch.manual_seed(1234)
    gen = Generator(h).to("cuda:0").train()
    x, s, n = (torch.randn(B, h.num_wv_feat, T, device="cuda"), torch.randn(B, h.spk_dim, device="cuda"),
               torch.randn(B, h.noise_dim, device="cuda"))
    out = {"config": f"cfg2 B={B} T={T} ResBlock2 train() under no_grad"}
    with torch.no_grad():
        for prec in ("auto", "bf16", "fp32"):
            gen.precision = prec
            ms = timed(lambda: gen(x, s, n), 20)
            prec = gen.train_precision
            out[f"native_{prec}_ms"] = ms
            out[f"native_{prec}_launches"] = gen.engine().last_launches
        out["delegate_fp32_ms"] = timed(lambda: gen._forward_autograd(x, s, n), 10)
        gen.eval()
        gen.precision = "bf16"
        t0 = time.perf_counter()
        gen(x, s, n)
        torch.cuda.synchronize()
        out["first_eval_after_train_ms (re-fold)"] = (time.perf_counter() - t0) * 1e3
        out["eval_bf16_ms"] = timed(lambda: gen(x, s, n), 20)
    out["audio_s_per_s_native_fp16"] = B * T / 50 / (out["native_fp16_ms"] * 1e-3)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
