import os, sys, pathlib
import pytest
ROOT = pathlib.Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (runs on the B200 box only)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu and os.environ.get("DB_DRY_RUN_GPU_TESTS") != "1":
        skip = pytest.mark.skip(reason="no CUDA device")
        for item in items:
            if "gpu" in item.keywords:
                item.add_marker(skip)


@pytest.fixture(autouse=True)
def _dry_run_gpu_tests(request):
    """DB_DRY_RUN_GPU_TESTS=1 (developer aid, GPU-less container): execute the GPU-marked tests through the kernel emulation, to catch
    mistakes in the GPU test files themselves before they reach the box.  Tests that need real CUDA objects still fail that way."""
    if os.environ.get("DB_DRY_RUN_GPU_TESTS") == "1" and "gpu" in request.keywords:
        sys.path.insert(0, str(ROOT / "tests"))
        from emu import emu_lib as E
        E.install()
        yield
        E.uninstall()
    else:
        yield


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    cache = {}
    def load(name):
        if name not in cache:
            cache[name] = np.load(GOLDEN / name, allow_pickle=False)
        return cache[name]
    return load
