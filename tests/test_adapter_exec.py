"""Executes the Chroma adapter (chroma_adapter/) -- not just type-checks it: tests/adapter_exec.cc is linked with the four
adapter sources, tests/mock_chroma (a functional stand-in for the slice of QDP++ they touch) and the real
libb200clover.so, registers the plugins in the factories, creates them by name from an XML-like parameter group, and
calls operator() with the CPU oracle as the caller's own linear operator.  A wrong pointer, stride, grid or parameter in
chroma_adapter/b200_clover_engine.h fails checkOperator or the residual re-check.  The 2-rank case forks one process per
T slab on cuda:0 and drives B200Glue::allgather / barrier through the mock's cross-process QDPInternal::globalSumArray.
Twin of lib/actions/ferm/invert/quda_solvers/syssolver_linop_clover_quda_w.h:71-648."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ADAPTER = os.path.join(ROOT, "chroma_adapter")
SOURCES = ["syssolver_b200_clover_params.cc", "syssolver_linop_clover_b200_w.cc", "syssolver_mdagm_clover_b200_w.cc",
           "multi_syssolver_mdagm_clover_b200_w.cc"]


@pytest.fixture(scope="module")
def binary(tmp_path_factory):
    from chroma_b200 import build as b
    from oracle import oracle as orc
    b.build()
    orc.build()
    top = tmp_path_factory.mktemp("adapter_exec")
    dst = top / "actions" / "ferm" / "invert" / "b200_solvers"
    dst.mkdir(parents=True)
    for f in os.listdir(ADAPTER):
        if f.endswith((".h", ".cc")):
            shutil.copy(os.path.join(ADAPTER, f), dst / f)
    exe = str(top / "adapter_exec")
    libdir, orcdir = os.path.join(ROOT, "chroma_b200"), os.path.join(ROOT, "oracle")
    cmd = ["g++", "-std=c++11", "-O1", "-Wall", "-Wno-unused", "-o", exe,
           "-I", str(top), "-I", os.path.join(ROOT, "tests", "mock_chroma"), "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "adapter_exec.cc"), os.path.join(ROOT, "tests", "mock_chroma", "mock_qdp.cc")]
    cmd += [str(dst / s) for s in SOURCES]
    cmd += ["-L", libdir, "-lb200clover", "-L", orcdir, "-loracle", "-Wl,-rpath," + libdir, "-Wl,-rpath," + orcdir,
            "-Wl,-rpath-link,/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64", "-fopenmp"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-6000:]
    return exe


def test_adapter_exec_builds_and_links(binary):
    """CPU: the adapter, the mock and the engine library link into one executable (every ABI symbol the adapter uses resolves)."""
    assert os.path.exists(binary)


def run(binary, latt, nranks):
    env = dict(os.environ, OMP_NUM_THREADS="2", B200_PEER_TIMEOUT_S="120")
    r = subprocess.run([binary] + [str(x) for x in latt] + [str(nranks)], capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0 and "FAIL" not in r.stdout and "ADAPTER_EXEC PASSED" in r.stdout, r.stdout[-6000:] + r.stderr[-3000:]
    return r.stdout


@pytest.mark.gpu
def test_adapter_exec_single_rank(binary):
    out = run(binary, (4, 4, 4, 8), 1)
    assert "checkOperator aborts" in out and "RNG state untouched" in out


@pytest.mark.gpu
def test_adapter_exec_two_ranks_one_device(binary):
    """T split over two processes on cuda:0: Layout::logicalSize / nodeCoord -> b200_create, B200Glue comm callbacks -> IPC
    bootstrap, local links of each slab, halos and cross-rank sums inside the solves."""
    if os.environ.get("B200_SKIP_ONE_DEVICE_TESTS"):
        pytest.skip("disabled by B200_SKIP_ONE_DEVICE_TESTS")
    run(binary, (4, 4, 4, 8), 2)
