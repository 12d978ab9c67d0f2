"""Timeline of the C = 128 fused stage kernel (developer build -DV2W_TRACE_TC [-DV2W_TRTC_TILE=n]): one cfg5-per-GPU
forward; CTA 0's MMA warp and epilogue warp 4 print SM-clock stamps for tile n (tags: 1/2 input tile wait, 10+10s /
11+10s segment s issue start / end, 100+10j.. per-branch pass wait / start / end, 200/201/202 final pass)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from wavthruvec_pytorch_b200.build import build_library
build_library(force=True)
import torch
from bench import CONFIGS
from wavthruvec_pytorch_b200 import Generator, HParams
batch, frames, feat = CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "cfg5_per_gpu"]
h = HParams(); h.num_wv_feat = feat
torch.manual_seed(1234)
gen = Generator(h).to("cuda:0").eval()
x, s, n = torch.randn(batch, feat, frames, device="cuda"), torch.randn(batch, 192, device="cuda"), torch.randn(batch, 192, device="cuda")
with torch.no_grad():
    gen(x, s, n)
torch.cuda.synchronize()
