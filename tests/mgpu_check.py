#!/usr/bin/env python
"""Multi-GPU parity check, one rank per GPU (launch with torchrun --nproc-per-node N).

Every rank builds the same seeded GLOBAL fields, keeps its sub-lattice (T slabs, or a Pz x Pt grid of T x Z boxes with
MGPU_GRID=pz,pt), runs the engine with NVLink halo exchange / cross-GPU reductions, and compares its part of the result
with the CPU oracle applied to the global lattice:
Dslash, M, M^dagger to 1e-13 per site; GPU-built clover term (needs the gauge ghost slices); CG and BiCGStab
iteration counts against the CPU restatement.  Exit code 0 = all ranks passed.
MGPU_ONE_DEVICE=1 puts every rank on cuda:0 (time-sliced contexts, peer memory through CUDA IPC on one device): slow,
but it exercises the whole multi-rank protocol on a single-GPU box.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bench import make_comm  # noqa: E402
from chroma_b200 import fields  # noqa: E402
from chroma_b200 import lib as L  # noqa: E402
from chroma_b200.solver import Context  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def rel_site_err(a, b):
    a = a.reshape(a.shape[0], -1)
    b = b.reshape(b.shape[0], -1)
    nb = np.linalg.norm(b, axis=1)
    return float((np.linalg.norm(a - b, axis=1) / np.maximum(nb, 1e-300)).max())


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    if os.environ.get("MGPU_ONE_DEVICE"):
        local_rank = 0
    torch.cuda.set_device(local_rank)
    latt = tuple(int(x) for x in os.environ.get("MGPU_LATT", "8,8,8,%d" % (4 * world)).split(","))
    prec = os.environ.get("MGPU_PREC", "double")
    recon = int(os.environ.get("MGPU_RECON", "18"))
    tol = 1e-13 if prec == "double" else 2e-6
    npdt = np.float64 if prec == "double" else np.float32
    pz, pt = (int(x) for x in os.environ.get("MGPU_GRID", "1,%d" % world).split(","))
    assert pz * pt == world
    grid, coord = (1, 1, pz, pt), (0, 0, rank % pz, rank // pz)       # rank = pt_coord * Pz + pz_coord
    lo, hi = fields.grid_box(latt, grid, coord)
    V = int(np.prod(latt))
    Vh = V // 2
    Vh_loc = Vh // world
    my_sites = fields.sub_lattice(np.arange(V), latt, lo, hi)          # global cb2 index of every local site, local cb2 order

    u = fields.apply_bc(latt, fields.weak_gauge(latt, seed=11), (1, 1, 1, -1))
    op = orc.Op(latt, u, 0.1, 1.0)
    comm = make_comm(dist, rank, world)
    ctx = Context(latt, prec=prec, device=local_rank, proc_grid=grid, proc_coord=coord, comm=comm)
    u_loc = np.stack([u[mu][my_sites] for mu in range(4)])
    ctx.load_gauge(u_loc.astype(npdt), t_boundary=-1, reconstruct=recon)
    ok = True

    def slab_cb(full, cb):   # rows of checkerboard cb of a full-lattice array that live in my sub-lattice
        return full[my_sites[cb * Vh_loc:(cb + 1) * Vh_loc]]

    def gather_odd(sol):     # assemble the global odd-checkerboard field from every rank's local solution
        parts, where = [None] * world, [None] * world
        dist.all_gather_object(parts, sol.astype(np.float64))
        dist.all_gather_object(where, my_sites[Vh_loc:])
        full = np.zeros((V, 4, 3, 2))
        for r in range(world):
            full[where[r]] = parts[r]
        return full

    def report(name, err, bound):
        nonlocal ok
        good = err < bound
        ok &= good
        print("[rank %d] %-34s err %.3e  (< %.1e) %s" % (rank, name, err, bound, "ok" if good else "FAIL"), flush=True)

    # ---- GPU-built clover (uses ghost link slices) vs restated build
    ctx.make_clover(4.1, 0.5, 0.5)
    clov, inv = ctx.get_clover()
    want_clov = op.clov[my_sites]
    report("make_clover vs restated", float(np.abs(clov - want_clov).max()), 1e-12 if prec == "double" else 1e-5)
    report("ldagdlinv vs restated", float(np.abs(inv - slab_cb(op.invclov, 0)).max()), 1e-11 if prec == "double" else 1e-5)

    # ---- hopping term and full operator
    psi = fields.gaussian_fermion(latt, seed=12)
    for isign in (+1, -1):
        for out_cb in (0, 1):
            want = slab_cb(op.dslash(psi, isign, out_cb), out_cb)
            got = ctx.dslash(slab_cb(psi, 1 - out_cb).astype(npdt), isign, out_cb)
            report("dslash isign=%+d cb=%d" % (isign, out_cb), rel_site_err(got.astype(np.float64), want), tol)
    podd = fields.gaussian_fermion(latt, seed=13, cb=1)
    for isign in (+1, -1):
        want = slab_cb(op.apply(podd, isign), 1)
        got = ctx.matpc(slab_cb(podd, 1).astype(npdt), isign)
        report("M isign=%+d" % isign, rel_site_err(got.astype(np.float64), want), 2 * tol)

    # ---- global sums
    f = ctx.field(slab_cb(podd, 1).astype(npdt))
    n2 = ctx.dev_norm2(f)
    report("norm2 (cross-GPU sum)", abs(n2 - np.sum(podd[Vh:] ** 2)) / np.sum(podd[Vh:] ** 2), 1e-12 if prec == "double" else 1e-6)

    # ---- solvers
    rsd = 1e-8 if prec == "double" else 1e-5
    for name, code, ref in (("CG", L.B200_SOLVER_CG, op.solve_cg), ("BICGSTAB", L.B200_SOLVER_BICGSTAB, op.solve_bicgstab)):
        _, n_ref, _, _ = ref(podd, np.zeros_like(podd), rsd, 2000)
        sol, info = ctx.invert(slab_cb(podd, 1).astype(npdt), None, solver=code, rsd=rsd, max_iter=2000)
        full = gather_odd(sol)   # gather the solution pieces to check the true residual with the CPU operator
        res = podd - op.apply(full, +1)
        rel = np.sqrt(np.sum(res[Vh:] ** 2) / np.sum(podd[Vh:] ** 2))
        good = info.converged == 1 and abs(info.n_count - n_ref) <= max(2, 0.05 * n_ref) and rel < 20 * rsd
        ok &= good
        print("[rank %d] %-8s iters %d (cpu %d) true rel resid %.3e reported %.3e %s" %
              (rank, name, info.n_count, n_ref, rel, info.rel_resid, "ok" if good else "FAIL"), flush=True)

    # ---- HMC-side normal-equation solve and mixed-precision reliable-update CG (predicated fp64 launches across ranks)
    gather_full = gather_odd

    _, n_ref, _ = op.solve_mdagm_cg(podd, np.zeros_like(podd), rsd, 2000)
    sol, info = ctx.invert_mdagm(slab_cb(podd, 1).astype(npdt), None, solver=L.B200_SOLVER_CG, rsd=rsd, max_iter=2000)
    full = gather_full(sol)
    res = podd - op.apply(op.apply(full, +1), -1)
    rel = np.sqrt(np.sum(res[Vh:] ** 2) / np.sum(podd[Vh:] ** 2))
    good = info.converged == 1 and abs(info.n_count - n_ref) <= max(2, 0.05 * n_ref) and rel < 50 * rsd
    ok &= good
    print("[rank %d] MdagM CG iters %d (cpu %d) true rel resid %.3e %s" % (rank, info.n_count, n_ref, rel, "ok" if good else "FAIL"), flush=True)
    if prec == "double":
        _, n_ref, nupd, _ = op.solve_reliable_cg(podd, np.zeros_like(podd), 1e-10, 0.1, 2000)
        sol, info = ctx.invert_reliable(slab_cb(podd, 1), None, rsd=1e-10, delta=0.1, max_iter=2000)
        full = gather_full(sol)
        res = podd - op.apply(full, +1)
        rel = np.sqrt(np.sum(res[Vh:] ** 2) / np.sum(podd[Vh:] ** 2))
        good = info.converged == 1 and abs(info.n_count - n_ref) <= max(3, 0.08 * n_ref) and rel < 2e-9
        ok &= good
        print("[rank %d] reliable CG iters %d (cpu %d, %d updates) true rel resid %.3e %s" %
              (rank, info.n_count, n_ref, nupd, rel, "ok" if good else "FAIL"), flush=True)

    # ---- batched (multi-RHS) kernels across the cut: batched halos, per-right-hand-side cross-GPU reductions
    nr = 5
    srcs = np.stack([fields.gaussian_fermion(latt, seed=300 + i, cb=1) for i in range(nr)])
    fin = ctx.mfield(nr, np.stack([slab_cb(srcs[i], 1) for i in range(nr)]).astype(npdt))
    fout = ctx.mfield(nr)
    for isign in (+1, -1):
        ctx.dev_matpc(fout, fin, isign)
        got = fout.download()
        err = max(rel_site_err(got[i].astype(np.float64), slab_cb(op.apply(srcs[i], isign), 1)) for i in range(nr))
        report("batched M isign=%+d (%d rhs)" % (isign, nr), err, 2 * tol)
    psi_b = ctx.mfield(nr)
    infos = ctx.dev_invert(psi_b, fin, solver=L.B200_SOLVER_BICGSTAB, rsd=rsd, max_iter=2000)
    sol_b = psi_b.download()
    for i in range(nr):
        _, n_ref, _, _ = op.solve_bicgstab(srcs[i], np.zeros_like(srcs[i]), rsd, 2000)
        full = gather_odd(sol_b[i])
        res = srcs[i] - op.apply(full, +1)
        rel = np.sqrt(np.sum(res[Vh:] ** 2) / np.sum(srcs[i][Vh:] ** 2))
        good = infos[i].converged == 1 and abs(infos[i].n_count - n_ref) <= max(2, 0.08 * n_ref) and rel < 20 * rsd
        ok &= good
        print("[rank %d] batched BiCGStab rhs %d iters %d (cpu %d) true rel resid %.3e %s" %
              (rank, i, infos[i].n_count, n_ref, rel, "ok" if good else "FAIL"), flush=True)

    # ---- section 8 (f3/f4) across the cut: reliable BiCGStab, multi-shift CG, symmetric preconditioning
    del fin, fout, psi_b
    if prec == "double":
        _, n_ref, nupd, _ = op.solve_reliable_bicgstab(podd, np.zeros_like(podd), 1e-10, 0.1, 2000)
        sol, info = ctx.invert_reliable_bicgstab(slab_cb(podd, 1), None, rsd=1e-10, delta=0.1, max_iter=2000)
        full = gather_full(sol)
        res = podd - op.apply(full, +1)
        rel = np.sqrt(np.sum(res[Vh:] ** 2) / np.sum(podd[Vh:] ** 2))
        good = info.converged == 1 and abs(info.n_count - n_ref) <= max(4, 0.15 * n_ref) and rel < 2e-9
        ok &= good
        print("[rank %d] reliable BiCGStab iters %d (cpu %d, %d/%d updates) true rel resid %.3e %s" %
              (rank, info.n_count, n_ref, info.n_updates, nupd, rel, "ok" if good else "FAIL"), flush=True)
    shifts = [0.001, 0.05, 0.7]
    ref_ms, n_ref, _ = op.solve_multishift(podd, shifts, rsd, 2000)
    sol_ms, infos = ctx.invert_multishift(slab_cb(podd, 1).astype(npdt), shifts, rsd, max_iter=2000)
    good = all(i.converged == 1 for i in infos) and abs(infos[0].n_count - n_ref) <= max(2, 0.05 * n_ref)
    worst = 0.0
    for s_, sh in enumerate(shifts):
        full = gather_odd(sol_ms[s_])
        res = podd - op.apply(op.apply(full, +1), -1) - sh * full
        worst = max(worst, np.sqrt(np.sum(res[Vh:] ** 2) / np.sum(podd[Vh:] ** 2)))
    good = good and worst < 50 * rsd
    ok &= good
    print("[rank %d] multi-shift CG iters %d (cpu %d) worst true rel resid %.3e %s" % (rank, infos[0].n_count, n_ref, worst, "ok" if good else "FAIL"), flush=True)
    ctx.set_preconditioning(True)
    op.set_symmetric(True)
    for isign in (+1, -1):
        want = slab_cb(op.apply(podd, isign), 1)
        got = ctx.matpc(slab_cb(podd, 1).astype(npdt), isign)
        report("symmetric M isign=%+d" % isign, rel_site_err(got.astype(np.float64), want), 2 * tol)
    _, n_ref, _, _ = op.solve_bicgstab(podd, np.zeros_like(podd), rsd, 2000)
    sol, info = ctx.invert(slab_cb(podd, 1).astype(npdt), None, solver=L.B200_SOLVER_BICGSTAB, rsd=rsd, max_iter=2000)
    full = gather_odd(sol)
    res = podd - op.apply(full, +1)
    rel = np.sqrt(np.sum(res[Vh:] ** 2) / np.sum(podd[Vh:] ** 2))
    good = info.converged == 1 and abs(info.n_count - n_ref) <= max(2, 0.08 * n_ref) and rel < 20 * rsd
    ok &= good
    print("[rank %d] symmetric BiCGStab iters %d (cpu %d) true rel resid %.3e %s" % (rank, info.n_count, n_ref, rel, "ok" if good else "FAIL"), flush=True)

    ctx.close()
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU_CHECK %s (world %d = %d x %d in Z x T, lattice %s, %s, recon %d)" % ("PASSED" if flag.item() else "FAILED", world, pz, pt, latt, prec, recon), flush=True)
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
