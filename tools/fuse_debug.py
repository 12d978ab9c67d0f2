import os, sys, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
from helpers import dropin_in_state, load_golden, t
from oracle import fixtures
name = sys.argv[1] if len(sys.argv) > 1 else "rb2_768_stress"
g = load_golden(name)
def run(mask):
    os.environ["V2W_FUSE_MASK"] = str(mask)
    gen = dropin_in_state(g, cache=False).to("cuda:0"); gen.precision = "bf16"
    with torch.no_grad():
        y = gen(t(g.x).cuda(), t(g.spk).cuda(), t(g.noise).cuda())
    torch.cuda.synchronize()
    return y.cpu().numpy()
base = run(0)
for mask in (4, 8, 16, 28):
    try:
        y = run(mask)
        print(f"mask {mask}: vs unfused {fixtures.snr_db(base, y):.1f} dB, vs ref {fixtures.snr_db(g.y, y):.1f} dB, nan={np.isnan(y).sum()}", flush=True)
    except Exception as e:
        print(f"mask {mask}: EXC {type(e).__name__} {str(e)[:200]}", flush=True); break
