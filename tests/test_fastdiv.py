"""CPU check of the division-free site decode used by the batched kernels (FastDiv, chroma_b200/csrc/common.cuh)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


@pytest.mark.skipif(not (os.path.exists(NVCC) or shutil.which("nvcc")), reason="needs nvcc")
def test_fast_division_is_exact(tmp_path):
    exe = str(tmp_path / "fastdiv_check")
    nvcc = NVCC if os.path.exists(NVCC) else shutil.which("nvcc")
    r = subprocess.run([nvcc, "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "fastdiv_check.cu")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "FASTDIV OK" in r.stdout, r.stdout + r.stderr
