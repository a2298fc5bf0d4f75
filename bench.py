#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 clover engine (contract: see the task prompt / DESIGN.md section 6).

A "step" is ONE CG ITERATION of the even-odd preconditioned Wilson-clover solve (InvCG2_a loop body,
lib/actions/ferm/invert/invcg2.cc:158-220): M p, |Mp|^2, M^dag(Mp), r -= a.., |r|^2, psi += a p, p = r + b p --
i.e. two applications of the fused clover Dslash operator plus the fused solver BLAS.  Workload: 48^3 x 96, fp64,
weak-field SU(3) gauge field, Mass 0.1, clovCoeff 1.0, antiperiodic T, Gaussian odd-checkerboard source (synthetic).

  value     GFLOP/s of K iterations with every field resident in HBM (Chroma's count: 7824 flop per odd site per
            iteration = 2*3792 + 240, invcg2.cc:67,101-220), CUDA events on the engine's stream, max over ranks.
  e2e       same metric through the plugin-facing C ABI call b200_invert() with HOST (pinned) chi / psi buffers:
            H2D of source + initial guess, M^dag chi, the solver preamble, K iterations, true-residual check, D2H of psi.
  roofline  dominant kernel = dslash_kernel<EPI_M> (chi = A_oo x - 1/4 D_oe t): algorithmic bytes (144+8G)*8 = 2304 B
            per odd site / its CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs.
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
    ap.add_argument("--solve", action="store_true", help="also run a full solve to 1e-8 and report time-to-solution")
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
        "config": workload_config(args, 1),
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
            "l2_policy": "working set per step (gauge+clover+vectors, >10 GB) exceeds the 126 MB L2; no flush needed"}


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

    # ---------------- leg 2: operator alone + per-kernel roofline (single GPU only: events around each kernel)
    roof = None
    dsl = None
    out_f = ctx.field()
    if world == 1:
        ctx.dev_time_matpc(out_f, chi_f, +1, 3)
        barrier()
        ms_a, ms_b = ctx.dev_time_matpc(out_f, chi_f, +1, max(5, args.steps))
        R = 8 if args.prec == "double" else 4
        G = args.recon
        bytes_a, bytes_b = (120 + 8 * G) * R, (144 + 8 * G) * R
        peak, how = peaks()
        achieved = bytes_b * Vh / (ms_b * 1e-3) * 1e-9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "dram_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("dslash_kernel_EPI_M_bytes_per_launch")
            except Exception:
                traffic = None
        roof = {"bound": "hbm", "kernel": "dslash_kernel<EPI_M> (chi = A_oo x - 1/4 D_oe t)", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "peak_source": how, "traffic": traffic,
                "algorithmic_bytes_per_launch": bytes_b * Vh, "ms_per_launch": ms_b,
                "other_kernels": {"dslash_kernel<EPI_AINV> (t = A_ee^-1 D_eo x)": {
                    "achieved": bytes_a * Vh / (ms_a * 1e-3) * 1e-9, "frac": bytes_a * Vh / (ms_a * 1e-3) * 1e-9 / peak, "ms_per_launch": ms_a}}}
        dsl = {"gflops": FLOP_M * Vh / ((ms_a + ms_b) * 1e-3) * 1e-9,
               "hbm_gbs": (bytes_a + bytes_b) * Vh / ((ms_a + ms_b) * 1e-3) * 1e-9,
               "frac_of_peak": (bytes_a + bytes_b) * Vh / ((ms_a + ms_b) * 1e-3) * 1e-9 / peak, "ms_per_apply": ms_a + ms_b}
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

    # ---------------- leg 3: end to end through the host-pointer ABI call (what the Chroma adapter calls)
    # rsd = 0 never converges (cp <= 0 is false), so exactly `steps` iterations run.
    psi_np[...] = 0
    ctx.invert(chi_np, psi_np, solver=solver, rsd=0.0, max_iter=max(1, args.warmup))   # warm the path once
    barrier()
    e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    psi_np[...] = 0
    e4.record(stream)
    info = L.SolveInfo()
    rc = ctx.lib.b200_invert(ctx.h, C.c_void_p(psi_np.ctypes.data), C.c_void_p(chi_np.ctypes.data), ctx.prec, solver, 0.0,
                             args.steps, C.byref(info))
    L.check(rc)
    e5.record(stream)
    barrier()
    ms_e2e = max_over_ranks(e4.elapsed_time(e5))
    e2e_value = flop_iter * Vh_global * args.steps / (ms_e2e * 1e-3) * 1e-9
    cb_bytes = Vh * 24 * (8 if args.prec == "double" else 4)
    e2e = {"value": e2e_value, "unit": "GFLOP/s", "h2d_bytes_per_step": 2 * cb_bytes * world / args.steps,
           "d2h_bytes_per_step": cb_bytes * world / args.steps, "ms_per_call": ms_e2e,
           "call": "b200_invert(host psi, host chi, max_iter=steps): H2D chi+psi0, M^dag chi, preamble, %d iterations, "
                   "true residual, D2H psi" % args.steps}
    # where the end-to-end time goes (untimed diagnostics, after the headline call): each stage between two events
    try:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        psi_d = ctx.field()
        barrier()
        ev[0].record(stream)
        chi_f.upload(chi_np); psi_d.upload(psi_np)
        ev[1].record(stream)
        ctx.dev_invert(psi_d, chi_f, solver=solver, rsd=0.0, max_iter=args.steps)
        ev[2].record(stream)
        L.check(ctx.lib.b200_mfield_download(ctx.h, psi_d.h, 0, C.c_void_p(psi_np.ctypes.data), ctx.prec))
        ev[3].record(stream)
        barrier()
        e2e["breakdown_ms"] = {"h2d_chi_psi0": ev[0].elapsed_time(ev[1]), "device_solve": ev[1].elapsed_time(ev[2]),
                               "d2h_psi": ev[2].elapsed_time(ev[3]), "iterations_only": ms_step * args.steps}
        del psi_d
    except Exception as e:  # noqa
        e2e["breakdown_ms"] = {"error": str(e)}
    # the same call on PAGEABLE host buffers (what QDP++ fields are): the engine bounces them through pinned double
    # buffers with a team of host threads (engine_impl.cuh::h2d / d2h)
    try:
        chi_pg, psi_pg = np.array(chi_np, copy=True), np.zeros_like(chi_np)
        for rep in range(2):      # the first call also pays the page faults of the freshly allocated arrays
            psi_pg[...] = 0
            barrier()
            e8, e9 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e8.record(stream)
            L.check(ctx.lib.b200_invert(ctx.h, C.c_void_p(psi_pg.ctypes.data), C.c_void_p(chi_pg.ctypes.data), ctx.prec, solver, 0.0,
                                        args.steps, C.byref(L.SolveInfo())))
            e9.record(stream)
            barrier()
            ms_pg = max_over_ranks(e8.elapsed_time(e9))
        e2e["pageable_host_buffers"] = {"ms_per_call": ms_pg, "value": flop_iter * Vh_global * args.steps / (ms_pg * 1e-3) * 1e-9}
        del chi_pg, psi_pg
    except Exception as e:  # noqa
        e2e["pageable_host_buffers"] = {"error": str(e)}
    t_clock_end = time.time()

    # ---------------- optional: a real solve to 1e-8 (time to solution)
    solve = None
    if args.solve:
        psi2 = ctx.field(np.zeros_like(chi_np))
        barrier()
        inf = ctx.dev_invert(psi2, chi_f, solver=solver, rsd=1e-8 if args.prec == "double" else 1e-6, max_iter=10000)
        barrier()
        solve = {"solver": args.solver, "seconds": max_over_ranks(inf.secs), "iterations": inf.n_count, "converged": bool(inf.converged),
                 "rel_resid": inf.rel_resid, "gflops": flop_iter * Vh_global * inf.n_count / max(inf.secs, 1e-9) * 1e-9}
        if args.prec == "double":
            # the same system by mixed-precision reliable-update CG (fp32 inner, fp64 outer; reliable_cg.cc)
            psi3 = ctx.field(np.zeros_like(chi_np))
            barrier()
            ctx.dev_invert_reliable(psi3, chi_f, rsd=1e-8, delta=0.1, max_iter=50)      # builds the fp32 twin (untimed)
            solve["mixed_precision_cg"] = []
            for delta in (0.1, 0.01):
                psi3.zero()
                barrier()
                inf = ctx.dev_invert_reliable(psi3, chi_f, rsd=1e-8, delta=delta, max_iter=10000)
                barrier()
                solve["mixed_precision_cg"].append({"delta": delta, "seconds": max_over_ranks(inf.secs), "iterations": inf.n_count,
                                                    "fp64_residual_replacements": inf.n_updates, "converged": bool(inf.converged),
                                                    "rel_resid": inf.rel_resid})

    clocks.stop()
    ck = clocks.summary(wall0, t_clock_end)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            cpu = cpu_arm(args.steps, args.warmup)
        except Exception as e:  # noqa
            cpu = {"error": str(e)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64" if args.prec == "double" else "f32", "data": "synthetic",
            "config": workload_config(args, world), "clocks": ck, "e2e": e2e, "gpu_launches": launches,
            "roofline": roof, "cpu_baseline": cpu, "clover_dslash": dsl, "multi_rhs": mrhs, "solve": solve,
            "setup_s": t_setup, "timed_wall_s": wall1 - wall0,
        }
        print(json.dumps(line))
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
