import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(params=["general", "fused"])
def nms_path(request):
    """Runs a GPU test once per NMS kernel path: the general three-launch path (library default) and the single-launch
    path of nms_fused.cu (segments of <= 4096 boxes).  Both must give identical results."""
    from object_detectors_b200 import _lib
    lib = _lib.load()
    lib.b200_debug_set_nms_path(1 if request.param == "general" else 0)
    yield request.param
    lib.b200_debug_set_nms_path(-1)
