"""GPU tests at BASELINE.json's full single-GPU sizes.

configs[1]  24^3x48 clover Dslash apply, fp64, one B200: compared with the oracle directly (it still runs in seconds there),
            <= 1e-13 per site (fp32 path <= 2e-6), hopping term and full operator, both signs.
configs[2]  32^3x64 EO-prec clover CG, fp64 and fp32 paths, and configs[3]'s lattice 48^3x96 (BiCGStab, one GPU): the
            oracle would need minutes, so the checks are the size-independent properties the reference's own tests rely on --
            <chi, M psi> = <M^dag chi, psi> (mainprogs/tests/t_precact_4d.cc:83-104), linearity, A_ee^-1 A_ee = 1, and the
            TRUE residual of the solution recomputed with an independent application of the operator
            (mainprogs/tests/symm_prec_tests.cc:210-249), plus bit-reproducibility of a repeated solve and agreement of
            the fp64 solve with the mixed-precision one.
Fields at the two large sizes are generated on the GPU with bench.py's seeded generators (numpy would take minutes).
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from chroma_b200 import fields  # noqa: E402
from chroma_b200 import lib as L  # noqa: E402
from chroma_b200.solver import Context  # noqa: E402

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    """global relative difference |a - b| / |b|"""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.sqrt(np.sum((a - b) ** 2) / np.sum(b ** 2)))


def rel_site_err(a, b):
    a = np.asarray(a, dtype=np.float64).reshape(a.shape[0], -1)
    b = np.asarray(b, dtype=np.float64).reshape(b.shape[0], -1)
    nb = np.linalg.norm(b, axis=1)
    m = nb > 0
    return float((np.linalg.norm(a - b, axis=1)[m] / nb[m]).max())


def test_config1_24x48_operator_vs_oracle(oracle):
    """BASELINE configs[1]: random (not smooth) SU(3) field, Mass 0.1, clovCoeff 1, antiperiodic T."""
    latt = (24, 24, 24, 48)
    u = fields.apply_bc(latt, fields.random_gauge(latt, seed=11))
    op = oracle.Op(latt, u, 0.1, 1.0)
    psi = fields.gaussian_fermion(latt, seed=12)
    for prec, npdt, tol in (("double", np.float64, 1e-13), ("single", np.float32, 2e-6)):
        ctx = Context(latt, prec=prec)
        ctx.load_gauge(u.astype(npdt), t_boundary=-1)
        ctx.make_clover(4.1, 0.5, 0.5)
        Vh = ctx.Vh
        for isign in (+1, -1):
            for out_cb in (0, 1):
                want = op.dslash(psi, isign, out_cb)[out_cb * Vh:(out_cb + 1) * Vh]
                src = psi[(1 - out_cb) * Vh:(2 - out_cb) * Vh].astype(npdt)
                got = ctx.dslash(src, isign, out_cb)
                assert rel_site_err(got, want) < tol, (prec, isign, out_cb)
            odd = np.zeros_like(psi)
            odd[Vh:] = psi[Vh:]
            want = op.apply(odd, isign)[Vh:]
            got = ctx.matpc(psi[Vh:].astype(npdt), isign)
            assert rel_site_err(got, want) < 2 * tol, (prec, isign)
        if prec == "double":      # GPU-built clover term at scale against the restated build
            clov, inv = ctx.get_clover()
            assert np.abs(clov - op.clov).max() < 1e-12
            assert np.abs(inv - op.invclov[:Vh]).max() < 1e-10
        ctx.close()


def _big_context(latt, prec):
    import torch
    import bench
    dev = torch.device("cuda", 0)
    ctx = Context(latt, prec=prec)
    u = bench.torch_weak_gauge(latt, 0, latt, 11, 0.2, dev)
    bench.apply_bc_local(u, latt, True)
    ctx.load_gauge(u if prec == "double" else u.astype(np.float32), t_boundary=-1)
    del u
    ctx.make_clover(4.1, 0.5, 0.5)
    tdt = torch.float64 if prec == "double" else torch.float32
    chi = bench.torch_gaussian_source(latt, 0, 12, dev, tdt).numpy()
    eta = bench.torch_gaussian_source(latt, 0, 13, dev, tdt).numpy()
    torch.cuda.empty_cache()
    return ctx, chi, eta


def _properties(ctx, chi, eta, prec):
    """gamma5-hermiticity, linearity and A^-1 A = 1 on device-resident fields; returns the fields for reuse."""
    eps = 1e-11 if prec == "double" else 2e-5
    fc, fe, m1, m2, t = ctx.field(chi), ctx.field(eta), ctx.field(), ctx.field(), ctx.field()
    ctx.dev_matpc(m1, fe, +1)                       # M eta
    ctx.dev_matpc(m2, fc, -1)                       # M^dag chi
    a, b = ctx.dev_inner(fc, m1), ctx.dev_inner(m2, fe)
    assert abs(a - b) < eps * abs(a), (a, b)
    # linearity: M (2 chi - 3 eta) = 2 M chi - 3 M eta
    comb = ctx.field((2.0 * chi - 3.0 * eta).astype(chi.dtype))
    ctx.dev_matpc(t, comb, +1)
    ctx.dev_matpc(m2, fc, +1)
    lhs = t.download(np.float64)
    rhs = 2.0 * m2.download(np.float64) - 3.0 * m1.download(np.float64)
    assert rel_site_err(lhs, rhs) < (1e-12 if prec == "double" else 5e-5)
    # A_ee^-1 A_ee = 1 on the even checkerboard
    ctx.dev_clover_apply(m1, fc, 0, False)
    ctx.dev_clover_apply(m2, m1, 0, True)
    assert rel_site_err(m2.download(np.float64), chi) < (1e-12 if prec == "double" else 2e-5)
    return fc, fe, m1, m2, t


def _true_rel_resid(ctx, psi_f, chi_f, scratch):
    """|chi - M psi| / |chi| with an operator application that is independent of the solver's own bookkeeping."""
    ctx.dev_matpc(scratch, psi_f, +1)
    r = chi_f.download(np.float64) - scratch.download(np.float64)
    c = chi_f.download(np.float64)
    return float(np.sqrt(np.sum(r * r) / np.sum(c * c)))


@pytest.mark.parametrize("prec,rsd", [("double", 1e-8), ("single", 1e-5)])
def test_config2_32x64_cg_properties(prec, rsd):
    """BASELINE configs[2]: 32^3x64 EO-prec clover CG, fp64 and fp32 paths."""
    latt = (32, 32, 32, 64)
    ctx, chi, eta = _big_context(latt, prec)
    fc, fe, m1, m2, t = _properties(ctx, chi, eta, prec)
    psi = ctx.field()
    info = ctx.dev_invert(psi, fc, solver=L.B200_SOLVER_CG, rsd=rsd, max_iter=5000)
    assert info.converged == 1 and 0 < info.n_count < 5000
    rel = _true_rel_resid(ctx, psi, fc, t)
    assert rel < (20 if prec == "double" else 50) * rsd, rel
    assert abs(info.rel_resid - rel) < 5e-2 * rel + (1e-14 if prec == "double" else 1e-6)
    # a repeated solve reproduces the solution bit for bit (deterministic reductions)
    first = psi.download()
    info2 = ctx.dev_invert(psi.zero(), fc, solver=L.B200_SOLVER_CG, rsd=rsd, max_iter=5000)
    assert info2.n_count == info.n_count and np.array_equal(psi.download(), first)
    if prec == "double":
        # the mixed-precision reliable-update CG reaches the same fp64 target in about as many iterations
        infm = ctx.dev_invert_reliable(psi.zero(), fc, rsd=rsd, delta=0.1, max_iter=5000)
        assert infm.converged == 1 and infm.n_updates >= 1 and abs(infm.n_count - info.n_count) <= 0.1 * info.n_count + 3
        assert _true_rel_resid(ctx, psi, fc, t) < 20 * rsd
        assert rel_err(psi.download(), first) < 1e-5
    ctx.close()


def test_config3_48x96_bicgstab_properties():
    """BASELINE configs[3]'s lattice on one GPU: 48^3x96 clover BiCGStab (the T-split runs are tests/test_multi_gpu.py and
    bench.py --gpus N), plus the multi-shift CG and the symmetric operator at this size."""
    latt = (48, 48, 48, 96)
    ctx, chi, eta = _big_context(latt, "double")
    fc, fe, m1, m2, t = _properties(ctx, chi, eta, "double")
    psi = ctx.field()
    info = ctx.dev_invert(psi, fc, solver=L.B200_SOLVER_BICGSTAB, rsd=1e-8, max_iter=5000)
    assert info.converged == 1 and 0 < info.n_count < 500
    rel = _true_rel_resid(ctx, psi, fc, t)
    assert rel < 1e-7 and abs(info.rel_resid - rel) < 5e-2 * rel
    sol = psi.download()
    # CG on the normal equations must land on the same solution
    infc = ctx.dev_invert(psi.zero(), fc, solver=L.B200_SOLVER_CG, rsd=1e-9, max_iter=5000)
    assert infc.converged == 1
    assert rel_err(psi.download(), sol) < 1e-5
    # mixed-precision BiCGStab: same target, comparable iteration count
    infr = ctx.dev_invert_reliable_bicgstab(psi.zero(), fc, rsd=1e-8, delta=0.1, max_iter=5000)
    assert infr.converged == 1 and infr.n_updates >= 1 and infr.n_count <= 1.3 * info.n_count + 4
    assert _true_rel_resid(ctx, psi, fc, t) < 1e-7
    # multi-shift: the smallest shift needs about as many iterations as CG on M^dag M; every shift meets its target
    shifts = [1e-3, 0.05, 1.0]
    mp = ctx.mfield(3)
    infos = ctx.dev_invert_multishift(mp, fc, shifts, 1e-8, max_iter=5000)
    assert all(i.converged == 1 and i.rel_resid < 1e-7 for i in infos)
    del mp
    # symmetric preconditioning at scale: hermiticity again, and S = A_oo^-1 M
    ctx.dev_matpc(m1, fe, +1)                        # M eta (asymmetric)
    ctx.set_preconditioning(True)
    ctx.dev_matpc(m2, fe, +1)                        # S eta
    ctx.dev_clover_apply(t, m1, 1, True)             # A_oo^-1 M eta
    assert rel_site_err(m2.download(), t.download()) < 1e-12
    ctx.dev_matpc(t, fc, -1)                         # S^dag chi
    a, b = ctx.dev_inner(fc, m2), ctx.dev_inner(t, fe)
    assert abs(a - b) < 1e-11 * abs(a)
    ctx.close()
