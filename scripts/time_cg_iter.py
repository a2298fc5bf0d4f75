#!/usr/bin/env python
"""A/B timing aid: ms per CG and BiCGStab iteration (48^3x96 fp64, device-resident, CUDA events) for the library selected
by B200_LIB_TAG.  Prints one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from chroma_b200 import lib as L  # noqa: E402
from chroma_b200.solver import Context  # noqa: E402

latt = tuple(int(x) for x in os.environ.get("PROF_LATT", "48,48,48,96").split(","))
dev = torch.device("cuda", 0)
ctx = Context(latt, prec="double")
u = bench.torch_weak_gauge(latt, 0, latt, 11, 0.2, dev)
bench.apply_bc_local(u, latt, True)
ctx.load_gauge(u, t_boundary=-1)
del u
ctx.make_clover(4.1, 0.5, 0.5)
chi = bench.torch_gaussian_source(latt, 0, 12, dev, torch.float64).numpy()
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
out = {"lib_tag": os.environ.get("B200_LIB_TAG", ""), "lattice": list(latt)}
chi_f, psi_f = ctx.field(chi), ctx.field()
for name, solver in (("cg", L.B200_SOLVER_CG), ("bicgstab", L.B200_SOLVER_BICGSTAB)):
    best = 1e9
    for rep in range(3):
        ctx.dev_iterate_begin(psi_f.zero(), chi_f, solver)
        ctx.dev_iterate(solver, 5)
        ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.dev_iterate(solver, 20)
        e1.record(stream)
        ctx.sync(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 20)
    out["ms_per_%s_iteration" % name] = best
info = ctx.dev_invert(psi_f.zero(), chi_f, solver=L.B200_SOLVER_CG, rsd=1e-8, max_iter=2000)
out["cg_solve"] = {"iterations": info.n_count, "secs": info.secs, "rel_resid": info.rel_resid}
print(json.dumps(out))
