"""Multi-GPU parity (T-split, NVLink peer-memory halos and reductions): spawns scripts/mgpu_check.py under torchrun.
Needs >= 2 GPUs on the box; skipped otherwise (the driver's 1-GPU tier skips it, `gpurun --gpus N` runs it)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def run_check(world, env_extra, port):
    env = dict(os.environ, **env_extra)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "scripts", "mgpu_check.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MGPU_CHECK PASSED" in r.stdout, r.stdout[-4000:] + r.stderr[-2000:]


@pytest.mark.parametrize("prec,recon", [("double", "18"), ("double", "12"), ("single", "18")])
def test_two_gpu_parity(prec, recon):
    if ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    run_check(2, {"MGPU_PREC": prec, "MGPU_RECON": recon, "MGPU_LATT": "8,8,8,8"}, 29521)


def test_two_gpu_thin_slabs():
    """Local T extent 2: no interior slice at all, every site is a boundary site."""
    if ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    run_check(2, {"MGPU_LATT": "8,4,4,4"}, 29522)


def test_four_gpu_parity():
    if ngpu() < 4:
        pytest.skip("needs 4 GPUs")
    run_check(4, {"MGPU_LATT": "8,8,4,16"}, 29523)


def test_eight_gpu_parity():
    if ngpu() < 8:
        pytest.skip("needs 8 GPUs")
    run_check(8, {"MGPU_LATT": "8,4,4,32"}, 29524)
