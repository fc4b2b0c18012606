"""Builds s2geometry_b200/libs2b.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

nvcc cross-compiles without a GPU, so this runs in the development container; the resulting .so travels to the
GPU box with the repo snapshot (it is git-ignored, not gpurun-ignored).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libs2b.so")

OBJ = os.path.join(HERE, "build", "obj")

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",            # the reference forbids FP contraction (_fp_contract_off.h); fma() is spelled out where wanted
    "-prec-div=true", "-prec-sqrt=true",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O2",
]
LINK_FLAGS = ["-cudart", "static", "-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a", "-ldl"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu") or f.endswith(".cc"))


def _headers():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))] + [
        os.path.join(HERE, "..", "include", "s2b_capi.h"), __file__]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in sources() + _headers())


def _compile(nvcc, src, obj, verbose):
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return src, r


def build(force=False, verbose=False):
    """One object per translation unit (compiled in parallel, rebuilt only when the unit or any header changed),
    then one link. Objects live in s2geometry_b200/build/obj (git- and gpurun-ignored); only libs2b.so ships."""
    if not force and not needs_build():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libs2b.so cannot be built (and there is no CPU fallback)")
    os.makedirs(OBJ, exist_ok=True)
    newest_header = max(os.path.getmtime(h) for h in _headers())
    jobs, objs = [], []
    for src in sources():
        obj = os.path.join(OBJ, os.path.basename(src) + ".o")
        objs.append(obj)
        if force or verbose or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), newest_header):
            jobs.append((src, obj))
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        for src, r in pool.map(lambda j: _compile(nvcc, j[0], j[1], verbose), jobs):
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError("nvcc failed on %s" % src)
            if verbose:
                sys.stderr.write(r.stdout + r.stderr)
    r = subprocess.run([nvcc] + LINK_FLAGS + objs + ["-o", LIB], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed linking libs2b.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
