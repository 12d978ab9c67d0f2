"""Timeline of the generic tap-GEMM kernel on the narrow layers (C = 16 / 32 / 64, many rows): stamps of CTA 0's
MMA warp and epilogue warps for one steady-state tile.  Needs the developer build next to the product library:
    nvcc ... -DV2W_TRACE_TC -o wavthruvec_pytorch_b200/libv2w_trace.so   (tools/trace_small.py loads it if present)
usage (GPU box): python tools/trace_small.py C k [rows] [batch]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
from wavthruvec_pytorch_b200 import native  # noqa: E402

trace_lib = ROOT / "wavthruvec_pytorch_b200" / "libv2w_trace.so"
if trace_lib.exists():
    native.LIB_PATH = trace_lib
import launch_floor  # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 16
k = int(sys.argv[2]) if len(sys.argv) > 2 else 11
rows = int(sys.argv[3]) if len(sys.argv) > 3 else 64000
batch = int(sys.argv[4]) if len(sys.argv) > 4 else 16
if len(sys.argv) > 5 and sys.argv[5] == "ups":
    # the epilogue of a transposed conv: affine, fp32 + 16-bit output, no residual; two column tiles (phases)
    print(launch_floor.case(batch, rows, C, 2 * C, C, k, 1, resid=False, f32=True, b16=True, reps=3, affine=True))
else:
    print(launch_floor.case(batch, rows, C, C, C, k, 1, resid=True, f32=True, b16=True, reps=3))
