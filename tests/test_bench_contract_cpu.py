"""bench.py's output contract, checked without a GPU: the reference arm prints exactly ONE JSON line on stdout (also
under torchrun, where rank 0 alone reports), and the product arm refuses to run without a CUDA device instead of
falling back to anything on the CPU."""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libs2ref.so")

needs_ref = pytest.mark.skipif(not os.path.exists(REF_LIB), reason="oracle/_ref not built (run __graft_entry__.build())")


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _check_reference_line(stdout):
    lines = stdout.splitlines()
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "points/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]
    # the reference arm runs the product arm's config: the same object, key for key (bench.workload_config)
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.workload_config(d["config"]["points_per_gpu"])
    return d


@needs_ref
def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "1", "--points", "4000000"], cwd=ROOT,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert _check_reference_line(r.stdout)["n_gpus"] == 1


@needs_ref
def test_reference_arm_under_torchrun_reports_from_rank_0_only():
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", str(_free_port()), "bench.py", "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "1", "--points", "4000000"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert _check_reference_line(r.stdout)["n_gpus"] == 2


def test_product_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([sys.executable, "bench.py", "--steps", "1", "--warmup", "1"], cwd=ROOT, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode != 0 and r.stdout.strip() == ""
    assert "CUDA device" in r.stderr
