import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def ref():
    """The reference oracle (oracle/_ref/libs2ref.so = unmodified reference sources). Test infrastructure only."""
    from oracle import reflib
    return reflib.load()


@pytest.fixture(scope="session")
def hostsim():
    """Test-only host compilation of the device headers (tests/hostsim)."""
    import ctypes
    here = os.path.join(ROOT, "tests", "hostsim")
    lib = os.path.join(here, "libhostsim.so")
    srcs = [os.path.join(here, f) for f in os.listdir(here) if f.endswith(".cc")]
    hdrs = [os.path.join(ROOT, "s2geometry_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "s2geometry_b200", "csrc"))]
    if not os.path.exists(lib) or any(os.path.getmtime(s) > os.path.getmtime(lib) for s in srcs + hdrs):
        subprocess.check_call(["bash", os.path.join(here, "build.sh")])
    return ctypes.CDLL(lib)
