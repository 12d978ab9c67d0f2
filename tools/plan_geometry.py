"""Launch geometry of every op of a plan (v2w_plan_get_launch).  usage (GPU box): python tools/plan_geometry.py [cfg2] [precision]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from bench import CONFIGS  # noqa: E402
from wavthruvec_pytorch_b200 import Generator, HParams  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
batch, frames, feat = CONFIGS[cfg]
h = HParams()
h.num_wv_feat = feat
gen = Generator(h).to("cuda:0").eval()
eng = gen.engine()
ops, geo = eng.ops(batch, frames, prec, "cuda:0"), eng.launches(batch, frames, prec, "cuda:0")
names = [n for n, _ in geo[0]._fields_]
print("op kind rows n_tile x n_tiles | " + " ".join(names))
for i, (o, d) in enumerate(zip(ops, geo)):
    print(i, o.kind, o.rows, f"{o.n_tile} x {o.n_tiles}", "|", " ".join(str(getattr(d, n)) for n in names))
