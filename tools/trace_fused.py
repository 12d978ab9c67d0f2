"""Developer tool: print the SM-clock timeline of the stacked fused kernels (CTA 0).

    V2W_NVCC_EXTRA=-DV2W_TRACE python wavthruvec_pytorch_b200/build.py --force   # -> copy to libv2w_trace.so
    V2W_LIB=wavthruvec_pytorch_b200/libv2w_trace.so python tools/trace_fused.py --config cfg2
"""
import argparse
import sys
import warnings
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
warnings.filterwarnings("ignore")


def main():
    import torch
    from bench import CONFIGS, make_hparams
    from wavthruvec_pytorch_b200 import Generator
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--resblock", default="2")
    ap.add_argument("--precision", default="bf16")
    args = ap.parse_args()
    batch, frames, feat = CONFIGS[args.config]
    h = make_hparams(args, feat)
    torch.manual_seed(1234)
    dev = torch.device("cuda:0")
    gen = Generator(h).eval().to(dev)
    gen.precision = args.precision
    x = torch.randn(batch, feat, frames, device=dev)
    s = torch.randn(batch, 192, device=dev)
    n = torch.randn(batch, 192, device=dev)
    with torch.no_grad():
        gen(x, s, n)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
