"""The C-ABI library loads and exports every symbol include/s2b_capi.h declares (no compute calls: no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "s2b_capi.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(s2b_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_all_declared_symbols():
    from s2geometry_b200 import build, _lib
    path = build.build()
    lib = ctypes.CDLL(path)
    syms = declared_symbols()
    assert len(syms) >= 10
    for s in syms:
        assert hasattr(lib, s), "libs2b.so does not export %s" % s
    # the Python binding table covers exactly the header
    assert sorted(_lib.SIGNATURES) == syms


def test_argument_errors_do_not_need_a_gpu():
    from s2geometry_b200 import _lib
    lib = _lib.load()
    assert lib.s2b_version().startswith(b"s2b")
    assert lib.s2b_device_count() >= 0
    # invalid level is rejected before any CUDA call
    assert lib.s2b_encode_points(None, 0, 31, None) == -1
    assert b"level" in lib.s2b_last_error()
    assert lib.s2b_encode_points(None, 0, 30, None) == 0  # empty batch is a no-op


def test_no_silent_cpu_fallback():
    """Without a CUDA device a compute call must fail loudly (S2B_ECUDA), never return CPU results."""
    import numpy as np
    from s2geometry_b200 import _lib, cellid
    lib = _lib.load()
    if lib.s2b_device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(_lib.S2BError):
        cellid.from_points(np.array([[1.0, 2.0, 3.0]]))
