#!/usr/bin/env python
"""Time the once-per-gauge-field setup path on one B200 (48^3x96 fp64 by default): gauge upload (H2D + on-GPU
transposition), clover build (field strength + makeClov), LDL^dagger inverse.  Wall clock around synchronous ABI calls."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from bench import torch_weak_gauge, apply_bc_local  # noqa: E402
from chroma_b200.solver import Context  # noqa: E402

latt = tuple(int(x) for x in os.environ.get("PROF_LATT", "48,48,48,96").split(","))
dev = torch.device("cuda", 0)
ctx = Context(latt, prec="double")
u = torch_weak_gauge(latt, 0, latt, 11, 0.2, dev)
apply_bc_local(u, latt, True)
out = {"lattice": list(latt), "lib_tag": os.environ.get("B200_LIB_TAG", "")}
for rep in range(3):
    t0 = time.time(); ctx.load_gauge(u, t_boundary=-1); t1 = time.time()
    ctx.make_clover(4.1, 0.5, 0.5); t2 = time.time()
    out["load_gauge_ms"] = (t1 - t0) * 1e3
    out["make_clover_total_ms"] = (t2 - t1) * 1e3       # 2 x make_clover_kernel + D2D copy + ldagdlinv_kernel
# one checkerboard fermion (1 GB at 48^3x96) from / to PAGEABLE host memory: what a solve on QDP++ fields pays around the solver
chi = np.random.default_rng(1).standard_normal((ctx.Vh, 4, 3, 2))
f = ctx.field()
import ctypes as C  # noqa: E402
from chroma_b200 import lib as L  # noqa: E402
back = np.zeros_like(chi)          # the destination is reused, like the psi field of a propagator loop: no first-touch faults in the timing
for rep in range(4):
    t0 = time.time(); f.upload(chi); ctx.sync(); t1 = time.time()
    L.check(ctx.lib.b200_mfield_download(ctx.h, f.h, 0, C.c_void_p(back.ctypes.data), ctx.prec)); t2 = time.time()
    out["field_upload_pageable_ms"] = (t1 - t0) * 1e3
    out["field_download_pageable_ms"] = (t2 - t1) * 1e3
assert np.array_equal(back, chi)
out["nt_copy_env"] = os.environ.get("B200_NT_COPY", "default(1)")
out["copy_threads_env"] = os.environ.get("B200_COPY_THREADS", "default")
ctx.set_preconditioning(True)
t0 = time.time(); ctx.make_clover(4.1, 0.5, 0.5); out["make_clover_total_symmetric_ms"] = (time.time() - t0) * 1e3
print(json.dumps(out))
