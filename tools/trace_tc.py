"""Timeline of the generic tap-GEMM kernel (developer build -DV2W_TRACE_TC): one launch of a k = 3, C = 256
conv with several tiles per CTA pair; prints the stamps of CTA 0's MMA warp and epilogue warp 4.
usage (GPU box): V2W_NVCC_EXTRA=-DV2W_TRACE_TC python tools/trace_tc.py [k] [rows]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
from wavthruvec_pytorch_b200.build import build_library  # noqa: E402

build_library(force=True)
import launch_floor  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 3
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
both = len(sys.argv) > 3
print(launch_floor.case(74, rows, 256, 256, 256, k, 1, resid=both, f32=both, b16=True, reps=3))
