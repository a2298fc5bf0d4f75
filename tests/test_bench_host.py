"""Host-side logic of bench.py that needs no GPU: the per-kernel roofline table, the N=1 solve records, the DRAM-traffic
lookup, and the reference arm under torchrun (rank 0 alone prints the line, with every host core -- the round-1 arm
inherited OMP_NUM_THREADS=1 from torchrun and was a 1-core baseline at N > 1)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_kernel_table_cg_bytes_and_shares():
    Vh, peak = 48 * 48 * 48 * 96 // 2, 6550.1
    kms = [1.8, 2.0, 1.8, 2.1, 0.85]
    tab = bench.kernel_table("CG", kms, 8, 18, Vh, peak)
    by = {k["kernel"]: k for k in tab}
    assert by["dslash_kernel<EPI_AINV>"]["launches_per_iteration"] == 2
    assert by["dslash_kernel<EPI_AINV>"]["algorithmic_bytes_per_site"] == (120 + 8 * 18) * 8 == 2112
    assert by["dslash_kernel<EPI_M_NORM>"]["algorithmic_bytes_per_site"] == 2304
    assert by["dslash_kernel<EPI_M_CG>"]["algorithmic_bytes_per_site"] == 2496
    assert by["cg_update_kernel"]["algorithmic_bytes_per_site"] == 960
    # what the fused loop moves per odd site and iteration: 2*2112 + 2304 + 2496 + 960 = 9984 B (the unfused algorithm of
    # DESIGN.md section 3, 2*4416 + 192*8 = 10368 B, writes and re-reads M^dag M p; EPI_M_CG never stores it)
    assert sum(k["algorithmic_bytes_per_site"] * k["launches_per_iteration"] for k in tab) == 9984
    assert abs(sum(k["share_of_iteration"] for k in tab) - 1.0) < 1e-12
    dom = max(tab, key=lambda k: k["share_of_iteration"])
    assert dom["kernel"] == "dslash_kernel<EPI_AINV>"
    assert dom["achieved"] == pytest.approx(2112 * Vh / 1.8e-3 * 1e-9)
    assert dom["frac"] == pytest.approx(dom["achieved"] / peak)


def test_kernel_table_bicgstab_has_seven_launches():
    tab = bench.kernel_table("BICGSTAB", [0.6, 1.8, 2.1, 0.7, 1.8, 2.1, 1.1], 8, 18, 1000, 6550.1)
    assert sum(k["launches_per_iteration"] for k in tab) == 7
    assert {k["kernel"] for k in tab} >= {"bicg_p_kernel", "dslash_kernel<EPI_M_DOTR0>", "dslash_kernel<EPI_M_DOTX>", "bicg_update_kernel"}


def test_kernel_table_batched_bytes():
    """nrhs > 1: spinor streams once per right-hand side, links + clover once per batch."""
    tab = bench.kernel_table("CG", [11.0, 12.5, 11.0, 14.0, 9.0], 8, 18, 1000, 6550.1, nrhs=12)
    by = {k["kernel"]: k for k in tab}
    op = (72 + 8 * 18) * 8                                  # 8 links + the clover block of a site: 1728 B
    assert by["dslash_kernel<EPI_AINV>"]["algorithmic_bytes_per_site"] == 12 * (2112 - op) + op
    assert by["dslash_kernel<EPI_M_NORM>"]["algorithmic_bytes_per_site"] == 12 * (2304 - op) + op
    assert by["dslash_kernel<EPI_M_CG>"]["algorithmic_bytes_per_site"] == 12 * (2496 - op) + op
    assert by["cg_update_kernel"]["algorithmic_bytes_per_site"] == 12 * 960
    # AINV + M per right-hand side = the 1248 B the bench's multi_rhs.clover_dslash leg quotes
    assert (by["dslash_kernel<EPI_AINV>"]["algorithmic_bytes_per_site"] + by["dslash_kernel<EPI_M_NORM>"]["algorithmic_bytes_per_site"]) / 12 == 1248


def test_expected_records_cover_the_driver_configurations():
    exp = bench.load_expected()

    class A:
        lattice, solver, prec, recon = [48, 48, 48, 96], "CG", "double", 18
    rec = exp[bench.expected_key(A)]
    assert rec["iterations"] == 57 and set(rec["checksums"]) == {"norm2_M_chi", "norm2_psi", "norm2_chi"}
    A.lattice, A.solver = [64, 64, 64, 128], "BICGSTAB"
    assert bench.expected_key(A) in exp            # BASELINE config 5


def test_dram_traffic_lookup_matches_the_workload_only():
    bench.LATTICE_OF_RUN, bench.PREC_OF_RUN, bench.RECON_OF_RUN = [48, 48, 48, 96], "double", 18
    t = bench.dram_traffic("dslash_kernel<EPI_AINV>")
    assert t is not None and 0.99 < t / (2112 * 48 * 48 * 48 * 96 // 2) < 1.06
    bench.LATTICE_OF_RUN = [32, 32, 32, 64]
    assert bench.dram_traffic("dslash_kernel<EPI_AINV>") is None
    bench.LATTICE_OF_RUN = None                   # a multi-GPU run: the capture is a 1-GPU one
    assert bench.dram_traffic("dslash_kernel<EPI_AINV>") is None
    bench.LATTICE_OF_RUN = [48, 48, 48, 96]


def test_reference_arm_under_torchrun_uses_all_cores():
    """world_size 2 (gloo-free: the arm needs no process group): rank 0 prints one JSON line whose cpu_baseline.cores is the
    number of cores the process may use although torchrun exports OMP_NUM_THREADS=1; rank 1 exits 0 silently; both arms
    share the config dict (same partition string as the B200 arm at this N)."""
    env = dict(os.environ)
    env.pop("OMP_NUM_THREADS", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29571", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    line = json.loads(lines[0])
    ncores = len(os.sched_getaffinity(0))
    assert line["impl"] == "reference" and line["cpu_baseline"]["cores"] == ncores > 1 or ncores == 1
    assert line["config"]["partition"] == "T-split x2" and "24x24x24x48" in line["config"]["cpu_reference_sample"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["unit"] == "GFLOP/s" and line["value"] > 0
