"""What bounds a transposed-conv launch of the generic kernel: one cfg5-per-GPU-sized shape at every super-tile
height and with either output switched off (CUDA events around single v2w_tapgemm launches).
usage (GPU box): python tools/ups_probe.py"""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tools"))
import launch_floor
B, R = 64, 32000
def run(tag, *a, **k):
    r = launch_floor.case(*a, **k)
    by = r['B'] * r['rows'] * (r['cin'] * 2 + (r['n'] * 4 if r['f32'] else 0) + (r['n'] * 2 if r['b16'] else 0) + (r['n'] * 4 if r['resid'] else 0))
    print(tag, r['us_med'], "us", round(by / r['us_med'] / 1e6, 2), "TB/s", flush=True)
for ms in ("0", "1", "2", "4"):
    os.environ["V2W_TC_MSUB"] = ms
    run(f"ups4 msub={ms} both", B, R, 32, 32, 32, 3, 1, f32=True, b16=True, affine=True)
os.environ["V2W_TC_MSUB"] = "0"
run("ups4 f32 only", B, R, 32, 32, 32, 3, 1, f32=True, b16=False, affine=True)
run("ups4 b16 only", B, R, 32, 32, 32, 3, 1, f32=False, b16=True, affine=True)
run("ups4 both no affine", B, R, 32, 32, 32, 3, 1, f32=True, b16=True)
run("ups4 k=1", B, R, 32, 32, 32, 1, 1, f32=True, b16=True, affine=True)
run("ups3 both", B, R // 2, 64, 64, 64, 3, 1, f32=True, b16=True, affine=True)
run("ups3 b16 only", B, R // 2, 64, 64, 64, 3, 1, f32=False, b16=True, affine=True)
run("ups3 f32 only", B, R // 2, 64, 64, 64, 3, 1, f32=True, b16=False, affine=True)
run("ups3 as n=128 two tiles", B, R // 4, 64, 128, 64, 3, 1, f32=True, b16=True, affine=True)
run("ups2-like n=256 tile 128", B, R // 8, 128, 256, 128, 2, 1, f32=True, b16=True, affine=True)
