import sys
import warnings
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

warnings.filterwarnings("ignore", category=FutureWarning)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real sm_100 GPU (run with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a box without a GPU: tests marked `gpu` are skipped, not failed."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (sm_100); none visible")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """Every test session works against a freshly-built in-tree libv2w.so (nvcc cross-compiles
    on the CPU box; on the GPU box the prebuilt file is up to date and this is a no-op)."""
    from wavthruvec_pytorch_b200.build import build_library
    build_library()
