#!/usr/bin/env python
"""Profiling driver: a few batched (multi-RHS) and single-RHS applications of M on a 32^3x64 lattice (for ncu)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from bench import torch_weak_gauge, apply_bc_local, torch_gaussian_source  # noqa: E402
from chroma_b200.solver import Context  # noqa: E402

latt = tuple(int(x) for x in os.environ.get("PROF_LATT", "32,32,32,64").split(","))
prec = os.environ.get("PROF_PREC", "double")
nr = int(os.environ.get("PROF_NRHS", "12"))
reps = int(os.environ.get("PROF_REPS", "3"))
dev = torch.device("cuda", 0)
ctx = Context(latt, prec=prec)
u = torch_weak_gauge(latt, 0, latt, 11, 0.2, dev)
apply_bc_local(u, latt, True)
ctx.load_gauge(u if prec == "double" else u.astype(np.float32), t_boundary=-1)
ctx.make_clover(4.1, 0.5, 0.5)
chi = torch_gaussian_source(latt, 0, 12, dev, torch.float64 if prec == "double" else torch.float32).numpy()
fin, fout = ctx.mfield(nr), ctx.mfield(nr)
for i in range(nr):
    fin.upload(np.roll(chi, 3 * i + 1, axis=0), i)
a, b = ctx.dev_time_matpc(fout, fin, +1, reps)
print("batched  nrhs=%d: AINV %.3f ms  M %.3f ms  per rhs %.3f ms" % (nr, a, b, (a + b) / nr))
f1, f2 = ctx.field(chi), ctx.field()
a1, b1 = ctx.dev_time_matpc(f2, f1, +1, reps)
print("single          : AINV %.3f ms  M %.3f ms  -> batched speed-up per rhs %.2fx" % (a1, b1, (a1 + b1) / ((a + b) / nr)))
ctx.close()
