import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_cuda() -> bool:
    try:
        import ctypes as C
        import rmcv_b200
        lib = rmcv_b200.load_library()
        n = C.c_int(0)
        lib.rmcv_device_count(C.byref(n))
        return n.value > 0
    except Exception:
        return False


_HAS_CUDA = None


def pytest_collection_modifyitems(config, items):
    global _HAS_CUDA
    gpu_items = [it for it in items if "gpu" in it.keywords]
    if not gpu_items:
        return
    if _HAS_CUDA is None:
        _HAS_CUDA = _has_cuda()
    if not _HAS_CUDA:
        skip = pytest.mark.skip(reason="no CUDA device in this container (GPU tests run under gpurun)")
        for it in gpu_items:
            it.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    import __graft_entry__ as g
    g.build()
    yield
