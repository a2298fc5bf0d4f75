"""Multi-GPU parity (T and T x Z process grids, NVLink peer-memory halos and reductions): spawns tests/mgpu_check.py
under torchrun.  The N-GPU cases need >= N GPUs on the box and are skipped otherwise (`gpurun --gpus N` runs them); the
`one_device` cases put every rank on cuda:0 (time-sliced contexts sharing memory through CUDA IPC), so the whole
multi-rank protocol -- face packing, arrival flags, ghost reads, corner links of the clover build, in-kernel cross-rank
reductions -- is also exercised on a single-GPU box."""
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
           "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_check.py")]
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


@pytest.mark.parametrize("grid,latt", [("2,1", "8,8,8,8"), ("2,2", "8,4,8,8")])
def test_txz_grid_parity(grid, latt):
    """Z-only and T x Z process grids, one rank per GPU."""
    world = int(grid.split(",")[0]) * int(grid.split(",")[1])
    if ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    run_check(world, {"MGPU_LATT": latt, "MGPU_GRID": grid}, 29525)


def test_eight_gpu_txz_grid():
    if ngpu() < 8:
        pytest.skip("needs 8 GPUs")
    run_check(8, {"MGPU_LATT": "8,4,8,16", "MGPU_GRID": "2,4"}, 29526)


@pytest.mark.parametrize("grid,latt,prec,recon", [("1,2", "4,4,4,8", "double", "18"), ("2,1", "4,4,8,4", "double", "18"),
                                                  ("2,2", "4,4,8,8", "double", "18"), ("2,2", "4,4,4,4", "double", "12"),
                                                  ("2,2", "4,4,8,8", "single", "18")])
def test_one_device_grid_parity(grid, latt, prec, recon):
    """Every rank on cuda:0.  (2,2) on 4^4 has local extent 2 in both split directions: no interior site at all."""
    if ngpu() < 1:
        pytest.skip("needs a GPU")
    if os.environ.get("B200_SKIP_ONE_DEVICE_TESTS"):
        pytest.skip("disabled by B200_SKIP_ONE_DEVICE_TESTS")
    world = int(grid.split(",")[0]) * int(grid.split(",")[1])
    run_check(world, {"MGPU_LATT": latt, "MGPU_GRID": grid, "MGPU_PREC": prec, "MGPU_RECON": recon, "MGPU_ONE_DEVICE": "1"}, 29527)


def test_one_device_split_reduction():
    """The fused single-launch split-lattice Dslash with the split reduction path (partials + one-CTA finish kernel that
    also does the cross-rank mailbox sum), forced on a small lattice by B200_SPLIT_MIN_BLOCKS=0."""
    if ngpu() < 1:
        pytest.skip("needs a GPU")
    if os.environ.get("B200_SKIP_ONE_DEVICE_TESTS"):
        pytest.skip("disabled by B200_SKIP_ONE_DEVICE_TESTS")
    run_check(2, {"MGPU_LATT": "4,4,4,8", "MGPU_GRID": "1,2", "MGPU_ONE_DEVICE": "1", "B200_SPLIT_MIN_BLOCKS": "0"}, 29528)
