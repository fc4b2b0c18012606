"""The C++ sugar header (include/s2b/s2b_batch.h) compiles against the C ABI everywhere and passes its checks on a GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_batch_api")


def build_exe():
    from s2geometry_b200 import build
    lib = build.build()
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = ["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"),
           os.path.join(ROOT, "tests", "cpp", "test_batch_api.cc"), "-o", EXE, lib, "-Wl,-rpath," + os.path.dirname(lib),
           "-L", os.path.join(cuda, "lib64"), "-lcudart", "-Wl,-rpath," + os.path.join(cuda, "lib64")]
    subprocess.check_call(cmd)
    return EXE


def test_cpp_header_compiles_and_links():
    assert os.path.exists(build_exe())


@pytest.mark.gpu
def test_cpp_batch_api_on_gpu():
    exe = build_exe()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all checks passed" in r.stdout
