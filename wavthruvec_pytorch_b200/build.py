"""In-tree build of libv2w.so (the C-ABI library of include/v2w.h) for sm_100a.

nvcc cross-compiles without a GPU, so this runs on the CPU build box; the resulting .so is
git-ignored but travels to the GPU box with the working tree.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
REPO_ROOT = PKG_DIR.parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libv2w.so"
SOURCES = [CSRC / "model.cpp", CSRC / "npy_io.cpp", CSRC / "plan.cu", CSRC / "mel.cu", CSRC / "pack_dev.cu"]
HEADERS = sorted(CSRC.glob("*.h")) + sorted(CSRC.glob("*.cuh")) + [REPO_ROOT / "include" / "v2w.h"]

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-cudart", "static",
]


def _nvcc() -> str:
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(cand).exists():
        raise RuntimeError("nvcc not found (set NVCC or add /usr/local/cuda/bin to PATH)")
    return cand


def needs_build() -> bool:
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    return any(p.stat().st_mtime > t for p in SOURCES + HEADERS)


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """Compile csrc/ into wavthruvec_pytorch_b200/libv2w.so; no-op when up to date."""
    if not force and not needs_build():
        return LIB_PATH
    extra = os.environ.get("V2W_NVCC_EXTRA", "").split()   # e.g. -DV2W_TRACE (developer builds)
    cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-I", str(REPO_ROOT / "include"), "-I", str(CSRC)]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    tmp = LIB_PATH.with_suffix(".so.tmp")
    cmd += ["-o", str(tmp), *map(str, SOURCES)]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    if verbose:
        sys.stderr.write(proc.stdout + proc.stderr)
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
