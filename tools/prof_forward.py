"""One eval()-mode forward of a BASELINE configuration between cudaProfilerStart/Stop, for
`ncu --profile-from-start off` (tools/gpu_r2.sh).  usage: prof_forward.py cfg3 [resblock] [precision]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from bench import CONFIGS  # noqa: E402
from wavthruvec_pytorch_b200 import Generator, HParams  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
batch, frames, feat = CONFIGS[cfg]
h = HParams()
h.num_wv_feat = feat
if len(sys.argv) > 2 and sys.argv[2] == "1":
    h.resblock = '1'
torch.manual_seed(1234)
gen = Generator(h).to("cuda:0").eval()
if len(sys.argv) > 3:
    gen.precision = sys.argv[3]
x, s, n = (torch.randn(batch, feat, frames, device="cuda"), torch.randn(batch, h.spk_dim, device="cuda"),
           torch.randn(batch, h.noise_dim, device="cuda"))
with torch.no_grad():
    for _ in range(3):
        y = gen(x, s, n)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    y = gen(x, s, n)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("ok", cfg, tuple(y.shape), gen.engine().last_launches, "launches")
