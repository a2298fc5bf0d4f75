#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 clover engine (contract: see the task prompt / DESIGN.md section 6).

A "step" is ONE CG ITERATION of the even-odd preconditioned Wilson-clover solve (InvCG2_a loop body,
lib/actions/ferm/invert/invcg2.cc:158-220): M p, |Mp|^2, M^dag(Mp), r -= a.., |r|^2, psi += a p, p = r + b p --
i.e. two applications of the fused clover Dslash operator plus the fused solver BLAS.  Workload: 48^3 x 96, fp64,
weak-field SU(3) gauge field, Mass 0.1, clovCoeff 1.0, antiperiodic T, Gaussian odd-checkerboard source (synthetic).

  value     GFLOP/s of K iterations with every field resident in HBM (Chroma's count: 7824 flop per odd site per
            iteration = 2*3792 + 240, invcg2.cc:67,101-220), CUDA events on the engine's stream, max over ranks.
  e2e       same metric through the plugin-facing C ABI call b200_invert() with PAGEABLE host chi / psi buffers (what
            QDP++ fields are): H2D of the source (the zero initial guess the propagator call site passes is set on the
            device, not copied), M^dag chi, the solver preamble, K iterations, true-residual check, D2H of psi.  The same call on pinned buffers is reported beside it (e2e.pinned_host_buffers).
  roofline  the kernels the CG loop really launches, each timed with CUDA events on the engine's stream inside the loop
            (b200_dev_time_solver_kernels): EPI_AINV (x2), EPI_M_NORM, EPI_M_CG, cg_update.  The dominant one is the
            kernel with the largest share of the iteration; achieved = its algorithmic bytes /
            its duration, against MEASURED_PEAKS.json hbm_gbs.
  solve     the BASELINE metric's second half: a real solve to 1e-8 at every N -- seconds, iterations, the TRUE relative
            residual, and two N-independent checksums (|M chi|^2, |psi|^2 as cross-rank sums) compared with the N=1
            values in tests/golden/bench_expected.json; a mismatch makes the run exit non-zero.
  cpu_baseline  the reference's own Dslash<double> (oracle/_ref, compiled from /root/reference) composed into the same
            CG iteration by the CPU restatement, all host cores, on a bounded 24^3x48 sample (rank 0, N=1 only).

`--impl reference` times that CPU arm alone under the same metric/unit/config.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

FLOP_CG, FLOP_M = 7824.0, 3792.0
METRIC = "EO-prec Wilson-clover CG iteration throughput (2 clover-Dslash M applies + fused BLAS; 7824 flop/odd site)"
SAMPLE_LATT = (24, 24, 24, 48)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--lattice", type=int, nargs=4, default=[48, 48, 48, 96])
    ap.add_argument("--prec", default="double", choices=["double", "single"])
    ap.add_argument("--recon", type=int, default=18, choices=[18, 12])
    ap.add_argument("--solver", default="CG", choices=["CG", "BICGSTAB"])
    ap.add_argument("--nrhs", type=int, default=12, help="right-hand sides of the batched (propagator) leg; 1 = skip it")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-solve", action="store_true", help="skip the full solve to 1e-8 (time to solution + N-independent checksums)")
    ap.add_argument("--no-fp32", action="store_true", help="skip the fp32 leg (per-kernel roofline of the single-precision engine)")
    ap.add_argument("--write-expected", action="store_true", help="record this run's solve checksums in tests/golden/bench_expected.json")
    ap.add_argument("--grid", type=int, nargs=2, default=None, metavar=("PZ", "PT"),
                    help="process grid in Z x T (PZ*PT = --gpus); default: T split only (1 x N)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------- CPU arm
def cpu_arm(steps, warmup):
    """Seconds per CG iteration of the CPU reference path on a bounded sample; returns the cpu_baseline dict."""
    from oracle import oracle as orc
    from chroma_b200 import fields
    # every core this process may use: torchrun exports OMP_NUM_THREADS=1, which would make this a 1-core baseline
    try:
        ncores = len(os.sched_getaffinity(0))
    except Exception:
        ncores = os.cpu_count() or 1
    orc.set_num_threads(ncores)
    latt = SAMPLE_LATT
    u = fields.apply_bc(latt, fields.weak_gauge(latt, seed=11))
    op = orc.Op(latt, u, 0.1, 1.0)
    chi = fields.gaussian_fermion(latt, seed=12, cb=1)
    n_timed = max(1, min(steps, 20))
    t_cg, t_m, kind = orc.cg_bench(op, chi, max(1, min(warmup, 3)), n_timed, use_reference_dslash=True)
    Vh = op.Vh
    return {
        "value": FLOP_CG * Vh / t_cg * 1e-9, "unit": "GFLOP/s", "cores": orc.num_threads(), "kind": kind,
        "sample": "%d CG iterations on a %dx%dx%dx%d sub-lattice (same operator, parameters and per-site work as the "
                  "48^3x96 workload; throughput is per-site so it carries over); hopping term = reference Dslash<double> "
                  "(OpenMP, %d threads), clover apply + BLAS = CPU restatement" % ((n_timed,) + latt + (orc.num_threads(),)),
        "clover_dslash_gflops": FLOP_M * Vh / t_m * 1e-9,
        "ms_per_cg_iteration_on_sample": t_cg * 1e3,
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.time()
    cb = cpu_arm(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "GFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_cg_iteration_on_sample"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.time() - t0,
    }
    print(json.dumps(line))


def workload_config(args, n):
    return {"workload": "%dx%dx%dx%d EO-prec Wilson-clover %s, %s, recon-%d, Mass=0.1 clovCoeff=1.0 antiperiodic-T, weak-field gauge"
            % (tuple(args.lattice) + (args.solver, "fp64" if args.prec == "double" else "fp32", args.recon)),
            "lattice": list(args.lattice),
            "partition": ("T-split x%d" % n) if not args.grid or args.grid[0] == 1 else "Z x T grid %d x %d" % tuple(args.grid),
            "l2_policy": "working set per step (gauge+clover+vectors, >10 GB) exceeds the 126 MB L2; no flush needed",
            "cpu_reference_sample": "the CPU reference arm (--impl reference, cpu_baseline) times the same operator and CG iteration on a "
                                    "%dx%dx%dx%d sub-lattice with every host core; GFLOP/s is per-site work x sites / time, so it carries over" % SAMPLE_LATT}


# ----------------------------------------------------------------------------------------------- clocks sampler
class Clocks:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.samples = []
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()

    def summary(self, t0, t1):
        sm, smax, reasons = [], 0.0, set()
        for ts, line in self.samples:
            if ts < t0 or ts > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                smax = max(smax, float(f[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- GPU arm
def torch_weak_gauge(latt_local, t0_global, latt_global, seed, eps, device, z0_global=0):
    """Smooth SU(3) field, generated on the GPU (same recipe as fields.weak_gauge: Gram-Schmidt of 1+eps*G).  The random
    stream is keyed by (seed, mu, GLOBAL time slice, checkerboard), so every rank count / process grid sees the same
    global field (a rank of a Z-split grid draws the whole slice and keeps its z range, which is contiguous in cb2 order)."""
    import torch
    V = int(np.prod(latt_local))
    Vh = V // 2
    lt = latt_local[3]
    s3h = Vh // lt
    row = (latt_local[0] // 2) * latt_local[1]
    s3h_g = row * latt_global[2]
    out = np.empty((4, V, 3, 3, 2), dtype=np.float64)
    eye = torch.eye(3, dtype=torch.complex128, device=device)
    gen = torch.Generator(device=device)
    for mu in range(4):
        g = torch.empty((V, 3, 3, 2), device=device, dtype=torch.float64)
        for cb in range(2):
            for t in range(lt):
                gen.manual_seed(((seed * 4 + mu) * 2 + cb) * 100003 + (t0_global + t))
                lo = cb * Vh + t * s3h
                g[lo:lo + s3h] = torch.randn((s3h_g, 3, 3, 2), generator=gen, device=device, dtype=torch.float64)[z0_global * row:z0_global * row + s3h]
        m = eye + eps * torch.view_as_complex(g)
        r0 = m[:, 0, :]
        r0 = r0 / torch.linalg.norm(r0, dim=-1, keepdim=True)
        r1 = m[:, 1, :]
        r1 = r1 - r0 * torch.sum(torch.conj(r0) * r1, dim=-1, keepdim=True)
        r1 = r1 / torch.linalg.norm(r1, dim=-1, keepdim=True)
        r2 = torch.conj(torch.linalg.cross(r0, r1, dim=-1))
        u = torch.view_as_real(torch.stack([r0, r1, r2], dim=1))
        out[mu] = u.cpu().numpy()
        del g, m, u, r0, r1, r2
    torch.cuda.empty_cache()
    return out


def torch_gaussian_source(latt_local, t0_global, seed, device, dtype, z0_global=0, lz_global=None):
    """Gaussian odd-checkerboard source keyed by (seed, GLOBAL time slice): identical for every rank count / grid."""
    import torch
    Vh = int(np.prod(latt_local)) // 2
    lt = latt_local[3]
    s3h = Vh // lt
    row = (latt_local[0] // 2) * latt_local[1]
    s3h_g = row * (lz_global or latt_local[2])
    gen = torch.Generator(device=device)
    out = torch.empty((Vh, 4, 3, 2), device=device, dtype=torch.float64)
    for t in range(lt):
        gen.manual_seed(seed * 100003 + 7 + (t0_global + t))
        out[t * s3h:(t + 1) * s3h] = torch.randn((s3h_g, 4, 3, 2), generator=gen, device=device, dtype=torch.float64)[z0_global * row:z0_global * row + s3h]
    return out.to(dtype).cpu().pin_memory()


def apply_bc_local(u, latt_local, is_last_rank):
    """Antiperiodic T: U_t *= -1 on the last GLOBAL time slice (held by the last rank)."""
    if not is_last_rank:
        return
    V = u.shape[1]
    Vh = V // 2
    s3h = Vh // latt_local[3]
    for cb in range(2):
        u[3, cb * Vh + (latt_local[3] - 1) * s3h: (cb + 1) * Vh] *= -1.0


def make_comm(dist, rank, world):
    """b200_comm backed by torch.distributed (one-time bootstrap only)."""
    from chroma_b200 import lib as L
    import torch

    def allgather(user, send, recv, nbytes):
        try:
            buf = torch.frombuffer((C.c_char * nbytes).from_address(send), dtype=torch.uint8).clone()
            outs = [torch.empty_like(buf) for _ in range(world)]
            dist.all_gather(outs, buf)
            flat = torch.cat(outs).numpy().tobytes()
            C.memmove(recv, flat, nbytes * world)
            return 0
        except Exception as e:  # noqa
            sys.stderr.write("allgather failed: %s\n" % e)
            return 1

    def barrier(user):
        dist.barrier()
        return 0

    comm = L.Comm()
    comm.rank, comm.size = rank, world
    comm._ag = L.ALLGATHER_FN(allgather)
    comm._ba = L.BARRIER_FN(barrier)
    comm.allgather, comm.barrier, comm.user = comm._ag, comm._ba, None
    return comm


EXPECTED = os.path.join(ROOT, "tests", "golden", "bench_expected.json")


def load_expected():
    try:
        return json.load(open(EXPECTED))
    except Exception:
        return {}


def expected_key(args):
    return "%dx%dx%dx%d %s %s recon%d" % (tuple(args.lattice) + (args.solver, args.prec, args.recon))


LATTICE_OF_RUN, PREC_OF_RUN, RECON_OF_RUN = [48, 48, 48, 96], "double", 18     # set by run_b200 from its arguments


def dram_traffic(kernel_name):
    """dram__bytes_read+write per launch of `kernel_name` from the newest `ncu --set full` capture of THIS build, if one was
    committed (profiles/dram_traffic.json: {"kernels": {name: bytes}, "build": ...}); None otherwise."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json")))
        if list(t.get("lattice", [])) != list(LATTICE_OF_RUN) or t.get("prec") != PREC_OF_RUN or t.get("recon") != RECON_OF_RUN:
            return None              # the capture was taken on another workload: no traffic figure for this run
        for k, v in t.get("kernels", {}).items():
            if k in kernel_name:
                return v
    except Exception:
        pass
    return None


def kernel_table(solver_name, kms, R, G, Vh, peak, nrhs=1):
    """Per-launch table of one solver iteration: algorithmic bytes per odd site (every array element once
    per pass, neighbour re-reads served by L2) and the measured duration of every launch, in launch order.  nrhs > 1: the
    batched kernels -- spinor streams once per right-hand side, links and clover once per batch; bytes per site for the batch."""
    op_a = op_m = (72 + 8 * G) * R          # operator data of one site: 8 links of G reals + one clover block of 72 reals
    ainv, m, mcg, upd = (120 + 8 * G) * R, (144 + 8 * G) * R, (168 + 8 * G) * R, 120 * R
    if nrhs > 1:
        ainv, m, mcg, upd = (ainv - op_a) * nrhs + op_a, (m - op_m) * nrhs + op_m, (mcg - op_m) * nrhs + op_m, upd * nrhs
    if solver_name == "CG":
        rows = [("dslash_kernel<EPI_AINV> t = A_ee^-1 D_eo p", ainv), ("dslash_kernel<EPI_M_NORM> mp = A_oo p - 1/4 D_oe t, |mp|^2", m),
                ("dslash_kernel<EPI_AINV> t = A_ee^-1 D_eo^dag mp", ainv), ("dslash_kernel<EPI_M_CG> r -= a(A_oo mp - 1/4 D_oe^dag t), |r|^2", mcg),
                ("cg_update_kernel psi += a p, p = r + b p", upd)]
    else:
        rows = [("bicg_p_kernel p = r + beta(p - omega v)", 96 * R * nrhs), ("dslash_kernel<EPI_AINV> t = A_ee^-1 D_eo p", ainv),
                ("dslash_kernel<EPI_M_DOTR0> v = M p, <r0|v>", mcg), ("bicg_s_kernel r -= alpha v", 72 * R * nrhs),
                ("dslash_kernel<EPI_AINV> t = A_ee^-1 D_eo r", ainv), ("dslash_kernel<EPI_M_DOTX> t = M r, <t|r>, |t|^2", m),
                ("bicg_update_kernel psi += omega r + alpha p, r -= omega t, |r|^2, <r0|r>", 192 * R * nrhs)]
    if len(kms) != len(rows):       # symmetric preconditioning adds a clover pass; not the default bench path
        rows = [("launch %d" % i, 0) for i in range(len(kms))]
    total = sum(kms)
    merged = {}
    for (name, nbytes), ms in zip(rows, kms):
        key = name.split(" ")[0]
        d = merged.setdefault(key, {"kernel": key, "what": [], "ms": [], "bytes": nbytes})
        d["what"].append(name); d["ms"].append(ms)
    out = []
    for key, d in merged.items():
        ms = sum(d["ms"]) / len(d["ms"])
        ach = d["bytes"] * Vh / (ms * 1e-3) * 1e-9 if ms > 0 else 0.0
        out.append({"kernel": key, "what": d["what"], "launches_per_iteration": len(d["ms"]), "ms_per_launch": ms,
                    "algorithmic_bytes_per_site": d["bytes"], "algorithmic_bytes_per_launch": d["bytes"] * Vh,
                    "achieved": ach, "frac": ach / peak, "share_of_iteration": sum(d["ms"]) / total if total > 0 else 0.0})
    return out


def fp32_leg(args, torch, dev, latt, solver, flop_iter, peak, barrier_main):
    """The CG loop on a single-precision engine (its own context: same gauge recipe, fp32 storage + arithmetic)."""
    from chroma_b200.solver import Context
    ctx = Context(latt, prec="single", device=dev.index)
    u = torch_weak_gauge(latt, 0, latt, 11, 0.2, dev)
    apply_bc_local(u, latt, True)
    ctx.load_gauge(u.astype(np.float32), t_boundary=-1, reconstruct=args.recon)
    del u
    ctx.make_clover(1.0 + 3.0 + 0.1, 0.5, 0.5)
    chi = torch_gaussian_source(latt, 0, 12, dev, torch.float32).numpy()
    chi_f, psi_f = ctx.field(chi), ctx.field(np.zeros_like(chi))
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    ctx.dev_iterate_begin(psi_f, chi_f, solver)
    ctx.dev_iterate(solver, args.warmup)
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    ctx.dev_iterate(solver, args.steps)
    e1.record(stream)
    ctx.sync()
    ms = e0.elapsed_time(e1) / args.steps
    kms = ctx.dev_time_solver_kernels(solver, max(5, args.steps))
    out = {"ms_per_step": ms, "gflops": flop_iter * ctx.Vh / (ms * 1e-3) * 1e-9,
           "kernels_in_loop": kernel_table(args.solver, kms, 4, args.recon, ctx.Vh, peak)}
    psi_f.zero()
    inf = ctx.dev_invert(psi_f, chi_f, solver=solver, rsd=1e-6, max_iter=10000)
    out["solve_to_1e-6"] = {"seconds": inf.secs, "iterations": inf.n_count, "rel_resid": inf.rel_resid, "converged": bool(inf.converged)}
    ctx.close()
    return out


def run_b200(args):
    import torch
    from chroma_b200 import lib as L
    from chroma_b200.solver import Context

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus %d must be launched with torchrun (one rank per GPU)" % args.gpus)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    comm = None
    if world > 1:
        import torch.distributed as dist
        # host-side bootstrap / result gathering only (gloo); the data path is NVLink peer memory inside the engine
        dist.init_process_group("gloo", rank=rank, world_size=world)
        comm = make_comm(dist, rank, world)

    global LATTICE_OF_RUN, PREC_OF_RUN, RECON_OF_RUN
    LATTICE_OF_RUN, PREC_OF_RUN, RECON_OF_RUN = (list(args.lattice) if world == 1 else None), args.prec, args.recon   # the capture is a 1-GPU one
    latt = tuple(args.lattice)
    pz, pt = tuple(args.grid) if args.grid else (1, world)
    assert pz * pt == world, "--grid PZ PT must multiply to the number of ranks"
    cz, ct = rank % pz, rank // pz                      # rank = pt_coord * PZ + pz_coord (include/b200_clover.h)
    assert latt[3] % pt == 0 and (latt[3] // pt) % 2 == 0, "T extent must split into even slabs"
    assert latt[2] % pz == 0 and (latt[2] // pz) % 2 == 0, "Z extent must split into even slabs"
    lt, lz = latt[3] // pt, latt[2] // pz
    latt_local = (latt[0], latt[1], lz, lt)
    solver = L.B200_SOLVER_CG if args.solver == "CG" else L.B200_SOLVER_BICGSTAB
    flop_iter = FLOP_CG if args.solver == "CG" else 2 * FLOP_M + 960.0

    t_setup = time.time()
    ctx = Context(latt, prec=args.prec, device=local_rank, proc_grid=(1, 1, pz, pt), proc_coord=(0, 0, cz, ct), comm=comm)
    u = torch_weak_gauge(latt_local, ct * lt, latt, 11, 0.2, dev, z0_global=cz * lz)
    apply_bc_local(u, latt_local, ct == pt - 1)
    ctx.load_gauge(u if args.prec == "double" else u.astype(np.float32), t_boundary=-1, reconstruct=args.recon)
    del u
    ctx.make_clover(1.0 + 3.0 + 0.1, 0.5, 0.5)
    Vh = ctx.Vh
    Vh_global = Vh * world
    npdt = np.float64 if args.prec == "double" else np.float32
    chi_host = torch_gaussian_source(latt_local, ct * lt, 12, dev, torch.float64 if args.prec == "double" else torch.float32,
                                     z0_global=cz * lz, lz_global=latt[2])
    psi_host = torch.zeros_like(chi_host).pin_memory()
    chi_np, psi_np = chi_host.numpy(), psi_host.numpy()
    t_setup = time.time() - t_setup

    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    clocks = Clocks(local_rank)
    clocks.start()

    # ---------------- leg 1: device-resident iterations
    chi_f, psi_f = ctx.field(chi_np), ctx.field(psi_np)
    ctx.dev_iterate_begin(psi_f, chi_f, solver)
    ctx.dev_iterate(solver, args.warmup)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = ctx.launch_count
    wall0 = time.time()
    e0.record(stream)
    ctx.dev_iterate(solver, args.steps)
    e1.record(stream)
    barrier()
    wall1 = time.time()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = ctx.launch_count - launches0
    ms_step = ms_total / args.steps
    value = flop_iter * Vh_global / (ms_step * 1e-3) * 1e-9

    # ---------------- leg 2: per-kernel roofline of the kernels the loop launches (CUDA events behind every launch of
    # the running CG / BiCGStab recurrences, b200_dev_time_solver_kernels), and the operator alone
    R = 8 if args.prec == "double" else 4
    G = args.recon
    peak, how = peaks()
    roof = None
    kern = None
    try:
        ctx.dev_time_solver_kernels(solver, 2)
        barrier()
        kms = ctx.dev_time_solver_kernels(solver, max(5, args.steps))
        kern = kernel_table(args.solver, kms, R, G, Vh, peak)
        dom = max(kern, key=lambda k: k["share_of_iteration"])
        roof = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved"], "peak": peak, "unit": "GB/s", "frac": dom["frac"],
                "peak_source": how, "traffic": dram_traffic(dom["kernel"]), "algorithmic_bytes_per_launch": dom["algorithmic_bytes_per_launch"],
                "ms_per_launch": dom["ms_per_launch"], "launches_per_iteration": dom["launches_per_iteration"],
                "share_of_iteration": dom["share_of_iteration"],
                "how": "CUDA events behind every launch of %d running %s iterations (engine stream)%s"
                       % (max(5, args.steps), args.solver, "" if world == 1 else "; rank 0 of %d, the kernel includes the halo wait" % world),
                "kernels_in_loop": kern}
    except Exception as e:  # noqa
        roof = {"error": str(e)}
    dsl = None
    out_f = ctx.field()
    if world == 1:
        ctx.dev_time_matpc(out_f, chi_f, +1, 3)
        barrier()
        ms_a, ms_b = ctx.dev_time_matpc(out_f, chi_f, +1, max(5, args.steps))
        bytes_a, bytes_b = (120 + 8 * G) * R, (144 + 8 * G) * R
        dsl = {"gflops": FLOP_M * Vh / ((ms_a + ms_b) * 1e-3) * 1e-9,
               "hbm_gbs": (bytes_a + bytes_b) * Vh / ((ms_a + ms_b) * 1e-3) * 1e-9,
               "frac_of_peak": (bytes_a + bytes_b) * Vh / ((ms_a + ms_b) * 1e-3) * 1e-9 / peak, "ms_per_apply": ms_a + ms_b,
               "kernels": {"EPI_AINV": {"ms": ms_a, "frac": bytes_a * Vh / (ms_a * 1e-3) * 1e-9 / peak},
                           "EPI_M": {"ms": ms_b, "frac": bytes_b * Vh / (ms_b * 1e-3) * 1e-9 / peak}}}
    else:
        ctx.dev_matpc(out_f, chi_f, +1)
        barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record(stream)
        for _ in range(args.steps):
            ctx.dev_matpc(out_f, chi_f, +1)
        e3.record(stream)
        barrier()
        ms_m = max_over_ranks(e2.elapsed_time(e3)) / args.steps
        dsl = {"gflops": FLOP_M * Vh_global / (ms_m * 1e-3) * 1e-9, "ms_per_apply": ms_m}

    # ---------------- leg 2b: the same iteration for a batch of right-hand sides (the 12 spin-colour sources of a propagator,
    # quarkprop4_w.cc:70-117) through the multi-RHS kernels: links and clover cross HBM once per batch
    mrhs = None
    if args.nrhs > 1:
        try:
            nr = args.nrhs
            chi_b, psi_b, out_b = ctx.mfield(nr), ctx.mfield(nr), ctx.mfield(nr)
            for i in range(nr):
                chi_b.upload(np.roll(chi_np, 7 * i + 1, axis=0), i)
            ctx.dev_iterate_begin(psi_b, chi_b, solver)
            ctx.dev_iterate(solver, args.warmup)
            barrier()
            e6, e7 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e6.record(stream)
            ctx.dev_iterate(solver, args.steps)
            e7.record(stream)
            barrier()
            ms_b = max_over_ranks(e6.elapsed_time(e7)) / args.steps
            R = 8 if args.prec == "double" else 4
            G = args.recon
            mrhs = {"nrhs": nr, "ms_per_iteration": ms_b, "ms_per_iteration_per_rhs": ms_b / nr,
                    "gflops": flop_iter * Vh_global * nr / (ms_b * 1e-3) * 1e-9,
                    "speedup_per_rhs_vs_single": ms_step / (ms_b / nr)}
            try:        # the batched iteration kernel by kernel (same events-behind-every-launch timing as the single-RHS table)
                kms_b = ctx.dev_time_solver_kernels(solver, max(3, args.steps // 2))
                pk, _ = peaks()
                mrhs["kernels_in_loop"] = [{k: v for k, v in row.items() if k != "what"} for row in
                                           kernel_table(args.solver, kms_b, R, G, Vh, pk, nr)]
                for row in mrhs["kernels_in_loop"]:
                    row["kernel"] = row["kernel"].replace("dslash_kernel", "dslash_mrhs_kernel")
            except Exception as e:  # noqa
                mrhs["kernels_in_loop"] = {"error": str(e)}
            if world == 1:
                ctx.dev_time_matpc(out_b, chi_b, +1, 2)
                barrier()
                mb_a, mb_b = ctx.dev_time_matpc(out_b, chi_b, +1, max(5, args.steps // 2))
                bytes_m = ((120.0 + (144 + 16 * G) / nr) * R) * nr          # algorithmic bytes per site for the whole batch
                peak, _ = peaks()
                mrhs["clover_dslash"] = {"ms_per_apply": mb_a + mb_b, "gflops": FLOP_M * Vh * nr / ((mb_a + mb_b) * 1e-3) * 1e-9,
                                         "algorithmic_bytes_per_site_per_rhs": bytes_m / nr,
                                         "hbm_gbs": bytes_m * Vh / ((mb_a + mb_b) * 1e-3) * 1e-9,
                                         "frac_of_peak": bytes_m * Vh / ((mb_a + mb_b) * 1e-3) * 1e-9 / peak}
            del chi_b, psi_b, out_b
        except Exception as e:  # noqa
            mrhs = {"error": str(e)}

    # ---------------- leg 3: end to end through the host-pointer ABI call (what the Chroma adapter calls) on PAGEABLE host
    # buffers -- QDP++ fields are ordinary heap memory; the engine bounces them through pinned double buffers with a
    # team of host threads (engine_impl.cuh::h2d / d2h).  rsd = 0 never converges (cp <= 0 is false), so exactly `steps`
    # iterations run.  The same call on pinned buffers is the side note.
    def timed_invert(chi_a, psi_a):
        psi_a[...] = 0
        barrier()
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record(stream)
        L.check(ctx.lib.b200_invert(ctx.h, C.c_void_p(psi_a.ctypes.data), C.c_void_p(chi_a.ctypes.data), ctx.prec, solver, 0.0,
                                    args.steps, C.byref(L.SolveInfo())))
        eb.record(stream)
        barrier()
        return max_over_ranks(ea.elapsed_time(eb))

    chi_pg, psi_pg = np.array(chi_np, copy=True), np.zeros_like(chi_np)       # plain numpy = pageable
    ctx.invert(chi_pg, psi_pg, solver=solver, rsd=0.0, max_iter=max(1, args.warmup))   # warm the path (and fault the pages in)
    timed_invert(chi_pg, psi_pg)
    ms_e2e = timed_invert(chi_pg, psi_pg)
    cb_bytes = Vh * 24 * (8 if args.prec == "double" else 4)
    e2e = {"value": flop_iter * Vh_global * args.steps / (ms_e2e * 1e-3) * 1e-9, "unit": "GFLOP/s",
           "h2d_bytes_per_step": cb_bytes * world / args.steps, "d2h_bytes_per_step": cb_bytes * world / args.steps,
           "ms_per_call": ms_e2e, "host_buffers": "pageable",
           "call": "b200_invert(host psi, host chi, max_iter=steps) on pageable numpy buffers: H2D chi (the initial guess psi0 = 0, as "
                   "quarkprop4_w.cc:74 passes it, is recognised by a host scan and set on the device, not copied), M^dag chi, preamble, "
                   "%d iterations, true residual, D2H psi" % args.steps}
    del chi_pg, psi_pg
    try:
        timed_invert(chi_np, psi_np)
        ms_pin = timed_invert(chi_np, psi_np)
        e2e["pinned_host_buffers"] = {"ms_per_call": ms_pin, "value": flop_iter * Vh_global * args.steps / (ms_pin * 1e-3) * 1e-9}
    except Exception as e:  # noqa
        e2e["pinned_host_buffers"] = {"error": str(e)}
    # where the end-to-end time goes (untimed diagnostics, after the headline call): each stage between two events
    try:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        psi_d = ctx.field()
        psi_np[...] = 0
        barrier()
        ev[0].record(stream)
        chi_f.upload(chi_np); psi_d.upload(psi_np)
        ev[1].record(stream)
        ctx.dev_invert(psi_d, chi_f, solver=solver, rsd=0.0, max_iter=args.steps)
        ev[2].record(stream)
        L.check(ctx.lib.b200_mfield_download(ctx.h, psi_d.h, 0, C.c_void_p(psi_np.ctypes.data), ctx.prec))
        ev[3].record(stream)
        barrier()
        e2e["breakdown_ms_pinned"] = {"h2d_chi_psi0": ev[0].elapsed_time(ev[1]), "device_solve": ev[1].elapsed_time(ev[2]),
                                      "d2h_psi": ev[2].elapsed_time(ev[3]), "iterations_only": ms_step * args.steps}
        del psi_d
    except Exception as e:  # noqa
        e2e["breakdown_ms_pinned"] = {"error": str(e)}
    t_clock_end = time.time()

    # ---------------- the BASELINE metric's second half: CG/BiCGStab time to solution, with N-independent checksums
    solve = None
    solve_ok = True
    if not args.no_solve:
        rsd = 1e-8 if args.prec == "double" else 1e-6
        psi2 = ctx.field(np.zeros_like(chi_np))
        ctx.dev_invert(psi2, chi_f, solver=solver, rsd=rsd, max_iter=3)              # warm the solve path
        psi2.zero()
        barrier()
        inf = ctx.dev_invert(psi2, chi_f, solver=solver, rsd=rsd, max_iter=10000)
        barrier()
        ctx.dev_matpc(out_f, chi_f, +1)
        sums = {"norm2_M_chi": ctx.dev_norm2(out_f), "norm2_psi": ctx.dev_norm2(psi2), "norm2_chi": ctx.dev_norm2(chi_f)}
        solve = {"solver": args.solver, "rsd_target": rsd, "seconds": max_over_ranks(inf.secs), "iterations": inf.n_count,
                 "converged": bool(inf.converged), "rel_resid": inf.rel_resid,
                 "rel_resid_how": "|chi - M psi| / |chi| recomputed with one more application of M after the loop (syssolver_linop_cg.h:80-87)",
                 "gflops": flop_iter * Vh_global * inf.n_count / max(inf.secs, 1e-9) * 1e-9, "checksums": sums}
        key = expected_key(args)
        exp = load_expected().get(key)
        if args.write_expected and rank == 0 and world == 1:
            allx = load_expected()
            allx[key] = {"iterations": inf.n_count, "checksums": sums, "written_by": "bench.py --write-expected on 1 GPU"}
            json.dump(allx, open(EXPECTED, "w"), indent=1, sort_keys=True)
            exp = allx[key]
        if exp is None:
            solve["check"] = "no N=1 record for this configuration in tests/golden/bench_expected.json"
        else:
            bad = []
            if not inf.converged or not (inf.rel_resid < 20 * rsd):
                bad.append("not converged: true rel resid %.3e" % inf.rel_resid)
            if abs(inf.n_count - exp["iterations"]) > max(1, round(0.03 * exp["iterations"])):
                bad.append("iterations %d vs %d at N=1" % (inf.n_count, exp["iterations"]))
            tol = {"norm2_M_chi": 1e-10, "norm2_chi": 1e-10, "norm2_psi": 1e-6} if args.prec == "double" else \
                  {"norm2_M_chi": 1e-5, "norm2_chi": 1e-5, "norm2_psi": 1e-3}
            for k, v in sums.items():
                if abs(v - exp["checksums"][k]) > tol[k] * abs(exp["checksums"][k]):
                    bad.append("%s %.15e vs %.15e at N=1" % (k, v, exp["checksums"][k]))
            solve["iterations_n1"] = exp["iterations"]
            solve["check"] = "ok: converged, iteration count and checksums match the N=1 record" if not bad else "FAILED: " + "; ".join(bad)
            solve_ok = not bad
        if args.prec == "double" and args.solver == "CG":
            # the same system by mixed-precision reliable-update CG (fp32 inner, fp64 outer; reliable_cg.cc)
            try:
                psi3 = ctx.field(np.zeros_like(chi_np))
                barrier()
                ctx.dev_invert_reliable(psi3, chi_f, rsd=1e-8, delta=0.1, max_iter=50)      # builds the fp32 twin (untimed)
                psi3.zero()
                barrier()
                inf = ctx.dev_invert_reliable(psi3, chi_f, rsd=1e-8, delta=0.1, max_iter=10000)
                barrier()
                solve["mixed_precision_cg"] = {"delta": 0.1, "seconds": max_over_ranks(inf.secs), "iterations": inf.n_count,
                                               "fp64_residual_replacements": inf.n_updates, "converged": bool(inf.converged),
                                               "rel_resid": inf.rel_resid}
                del psi3
            except Exception as e:  # noqa
                solve["mixed_precision_cg"] = {"error": str(e)}
        del psi2

    clocks.stop()
    ck = clocks.summary(wall0, t_clock_end)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            cpu = cpu_arm(args.steps, args.warmup)
        except Exception as e:  # noqa
            cpu = {"error": str(e)}

    # ---------------- fp32 leg: the same loop on the single-precision engine (per-kernel roofline fractions)
    fp32 = None
    if world == 1 and args.prec == "double" and not args.no_fp32:
        try:
            fp32 = fp32_leg(args, torch, dev, latt, solver, flop_iter, peak, barrier)
        except Exception as e:  # noqa
            fp32 = {"error": str(e)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64" if args.prec == "double" else "f32", "data": "synthetic",
            "config": workload_config(args, world), "clocks": ck, "e2e": e2e, "gpu_launches": launches,
            "roofline": roof, "cpu_baseline": cpu, "clover_dslash": dsl, "multi_rhs": mrhs, "solve": solve, "fp32": fp32,
            "setup_s": t_setup, "timed_wall_s": wall1 - wall0,
        }
        print(json.dumps(line))
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()
    if not solve_ok:
        sys.stderr.write("bench.py: solve check failed: %s\n" % (solve or {}).get("check"))
        sys.exit(3)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
