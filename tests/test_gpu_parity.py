"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): <= 1e-13 relative per site in fp64, <= 1e-6 in fp32, for the hopping term, the clover
apply / inverse and the full even-odd operator; solvers reach the same target residual with iteration counts within a
few percent of the CPU restatement of InvCG2 / InvBiCGStab.  Procedure mirrors the reference's own differential tests
(other_libs/cpp_wilson_dslash/tests/testDslashFull.cc:70-97, mainprogs/tests/t_lwldslash_sse.cc:217-241,
mainprogs/tests/symm_prec_tests.cc:210-249).
"""
import numpy as np
import pytest

from chroma_b200 import fields
from chroma_b200 import lib as L
from chroma_b200.solver import (CloverFermActParams, AnisoParam, Context, LinOpSysSolverB200Clover,
                                MdagMMultiSysSolverB200Clover, SolverFailure, SysSolverB200CloverParams)

pytestmark = pytest.mark.gpu

TOL = {"double": 1e-13, "single": 1e-6}
NP = {"double": np.float64, "single": np.float32}
LATTICES = [(4, 4, 4, 8), (6, 4, 2, 4), (8, 8, 8, 8), (16, 8, 8, 12)]


def rel_site_err(a, b):
    a = np.asarray(a, dtype=np.float64).reshape(a.shape[0], -1)
    b = np.asarray(b, dtype=np.float64).reshape(b.shape[0], -1)
    nb = np.linalg.norm(b, axis=1)
    m = nb > 0
    return float((np.linalg.norm(a - b, axis=1)[m] / nb[m]).max())


def setup(oracle, latt, prec="double", gauge="random", recon=L.B200_RECONS_NONE, aniso=False, bc=(1, 1, 1, -1), gpu_clover=False):
    u = fields.random_gauge(latt, seed=11) if gauge == "random" else fields.weak_gauge(latt, seed=11)
    u = fields.apply_bc(latt, u, bc)
    an = dict(anisoP=True, t_dir=3, xi_0=2.464, nu=0.95) if aniso else {}
    cR, cT = (0.91, 1.07) if aniso else (1.0, 1.0)
    op = oracle.Op(latt, u, 0.1, cR, cT, **an)
    cp = CloverFermActParams(Mass=0.1, clovCoeffR=cR, clovCoeffT=cT, anisoParam=AnisoParam(**an))
    ctx = Context(latt, prec=prec)
    ctx.load_gauge(u.astype(NP[prec]) if prec == "single" else u, aniso_coeff=cp.ferm_coeffs(), t_boundary=bc[3], reconstruct=recon)
    if gpu_clover:
        ctx.make_clover(*cp.derived(), aniso=aniso, t_dir=3)
    else:
        ctx.load_clover(op.clov.copy(), op.invclov.copy())
    return u, op, ctx, cp


@pytest.mark.parametrize("prec", ["double", "single"])
@pytest.mark.parametrize("latt", LATTICES)
def test_dslash_parity(oracle, latt, prec):
    u, op, ctx, _ = setup(oracle, latt, prec)
    psi = fields.gaussian_fermion(latt, seed=12)
    Vh = ctx.Vh
    for isign in (+1, -1):
        for out_cb in (0, 1):
            want = op.dslash(psi, isign, out_cb)[out_cb * Vh:(out_cb + 1) * Vh]
            src = psi[(1 - out_cb) * Vh:(2 - out_cb) * Vh].astype(NP[prec])
            got = ctx.dslash(src, isign, out_cb)
            assert rel_site_err(got, want) < TOL[prec], (isign, out_cb)
    ctx.close()


@pytest.mark.parametrize("recon", [L.B200_RECONS_NONE, L.B200_RECONS_12])
@pytest.mark.parametrize("aniso", [False, True])
def test_dslash_parity_aniso_recon(oracle, recon, aniso):
    latt = (8, 4, 4, 8)
    u, op, ctx, _ = setup(oracle, latt, "double", recon=recon, aniso=aniso)
    psi = fields.gaussian_fermion(latt, seed=12)
    Vh = ctx.Vh
    for isign in (+1, -1):
        for out_cb in (0, 1):
            want = op.dslash(psi, isign, out_cb)[out_cb * Vh:(out_cb + 1) * Vh]
            got = ctx.dslash(psi[(1 - out_cb) * Vh:(2 - out_cb) * Vh], isign, out_cb)
            assert rel_site_err(got, want) < 2e-13, (isign, out_cb)
    ctx.close()


@pytest.mark.parametrize("prec", ["double", "single"])
def test_clover_apply_and_inverse_parity(oracle, prec):
    latt = (8, 4, 4, 8)
    u, op, ctx, _ = setup(oracle, latt, prec)
    g = oracle.Geom(latt)
    psi = fields.gaussian_fermion(latt, seed=14)
    Vh = ctx.Vh
    for cb in (0, 1):
        want = oracle.clover_apply(g, psi, op.clov, cb)[cb * Vh:(cb + 1) * Vh]
        got = ctx.clover_apply(psi[cb * Vh:(cb + 1) * Vh].astype(NP[prec]), cb, inverse=False)
        assert rel_site_err(got, want) < TOL[prec]
    want = oracle.clover_apply(g, psi, op.invclov, 0)[:Vh]
    got = ctx.clover_apply(psi[:Vh].astype(NP[prec]), 0, inverse=True)
    assert rel_site_err(got, want) < TOL[prec]
    with pytest.raises(L.B200Error):
        ctx.clover_apply(psi[Vh:].astype(NP[prec]), 1, inverse=True)     # only the cb-0 inverse exists
    ctx.close()


@pytest.mark.parametrize("prec", ["double", "single"])
@pytest.mark.parametrize("latt", LATTICES)
def test_matpc_parity(oracle, latt, prec):
    """M and M^dagger against EvenOddPrecCloverLinOp::operator() restated (eoprec_clover_linop_w.cc:142-187)."""
    u, op, ctx, _ = setup(oracle, latt, prec)
    psi = fields.gaussian_fermion(latt, seed=12, cb=1)
    Vh = ctx.Vh
    for isign in (+1, -1):
        want = op.apply(psi, isign)[Vh:]
        got = ctx.matpc(psi[Vh:].astype(NP[prec]), isign)
        # the operator has a 4.1-ish diagonal and O(8) hopping part: same per-site criterion
        assert rel_site_err(got, want) < 2 * TOL[prec], isign
    ctx.close()


def test_matpc_parity_aniso_recon12(oracle):
    latt = (8, 4, 4, 8)
    u, op, ctx, _ = setup(oracle, latt, "double", recon=L.B200_RECONS_12, aniso=True)
    psi = fields.gaussian_fermion(latt, seed=12, cb=1)
    Vh = ctx.Vh
    for isign in (+1, -1):
        assert rel_site_err(ctx.matpc(psi[Vh:], isign), op.apply(psi, isign)[Vh:]) < 3e-13
    ctx.close()


@pytest.mark.parametrize("recon", [L.B200_RECONS_NONE, L.B200_RECONS_12])
@pytest.mark.parametrize("aniso,latt", [(False, (6, 4, 4, 8)), (True, (6, 4, 4, 8)), (False, (6, 6, 2, 2))])
def test_gpu_built_clover_matches_restated_build(oracle, recon, aniso, latt):
    """b200_make_clover (field strength + makeClov + LDL^dagger inverse on the GPU) against the restated
    mesField / makeClov / ldagdlinv (mesfield.cc:44-74, clover_term_qdp_w.h:416-519, 636-815).  6x6x2x2 has 72 sites
    per checkerboard: the last CTA of the site x plane build is partly empty, and every direction wraps after 2 or 6."""
    u, op, ctx, cp = setup(oracle, latt, "double", recon=recon, aniso=aniso, gpu_clover=True)
    clov, inv = ctx.get_clover()
    assert np.abs(clov - op.clov).max() < 1e-13
    assert np.abs(inv - op.invclov[:ctx.Vh]).max() < 1e-12
    g = oracle.Geom(latt)
    _, trlog = oracle.ldagdlinv(g, op.clov, 0)
    assert abs(ctx.clover_logdet() - trlog[:ctx.Vh].sum()) < 1e-9 * abs(trlog.sum())
    ctx.close()


def test_device_fields_norm_inner_hermiticity(oracle):
    """Device-resident interface + <chi, M psi> = <M^dag chi, psi> on the GPU (t_precact_4d.cc:83-104)."""
    latt = (8, 8, 4, 4)
    u, op, ctx, _ = setup(oracle, latt, "double")
    Vh = ctx.Vh
    a = fields.gaussian_fermion(latt, seed=20)[Vh:]
    b = fields.gaussian_fermion(latt, seed=21)[Vh:]
    fa, fb, fm, fn = ctx.field(a), ctx.field(b), ctx.field(), ctx.field()
    assert np.array_equal(fa.download(), a)                     # AoS -> SoA -> AoS round trip is exact
    assert abs(ctx.dev_norm2(fa) - np.sum(a * a)) < 1e-12 * np.sum(a * a)
    ca, cb_ = a[..., 0] + 1j * a[..., 1], b[..., 0] + 1j * b[..., 1]
    assert abs(ctx.dev_inner(fa, fb) - np.vdot(ca, cb_)) < 1e-11 * abs(np.vdot(ca, cb_))
    ctx.dev_matpc(fm, fb, +1)
    ctx.dev_matpc(fn, fa, -1)
    lhs, rhs = ctx.dev_inner(fa, fm), ctx.dev_inner(fn, fb)
    assert abs(lhs - rhs) < 1e-11 * abs(lhs)
    ctx.close()


@pytest.mark.parametrize("solver", ["CG", "BICGSTAB"])
@pytest.mark.parametrize("prec,rsd", [("double", 1e-8), ("single", 1e-5)])
def test_solver_matches_cpu_restatement(oracle, solver, prec, rsd):
    """BASELINE configs[0] on the GPU: 8^4 weak field, Mass 0.1, clovCoeff 1, antiperiodic T.  Same target residual;
    iteration count within a few percent of InvCG2_a / InvBiCGStab_a restated on the CPU."""
    latt = (8, 8, 8, 8)
    u, op, ctx, cp = setup(oracle, latt, prec, gauge="weak")
    chi = fields.gaussian_fermion(latt, seed=12, cb=1)
    Vh = ctx.Vh
    if solver == "CG":
        psi_ref, n_ref, resid_ref, rel_ref = op.solve_cg(chi, np.zeros_like(chi), rsd, 2000)
        code = L.B200_SOLVER_CG
    else:
        psi_ref, n_ref, resid_ref, rel_ref = op.solve_bicgstab(chi, np.zeros_like(chi), rsd, 2000)
        code = L.B200_SOLVER_BICGSTAB
    psi, info = ctx.invert(chi[Vh:].astype(NP[prec]), None, solver=code, rsd=rsd, max_iter=2000)
    assert info.converged == 1
    assert abs(info.n_count - n_ref) <= max(2, 0.05 * n_ref), (info.n_count, n_ref)
    # true residual, recomputed with the CPU operator
    full = np.zeros_like(chi)
    full[Vh:] = psi
    r = chi - op.apply(full, +1)
    rel = np.sqrt(np.sum(r[Vh:] ** 2) / np.sum(chi[Vh:] ** 2))
    assert rel < (10 * rsd if prec == "double" else 50 * rsd)
    if prec == "double":
        assert abs(info.rel_resid - rel) < 1e-3 * rel + 1e-14
        assert rel_site_err(psi, psi_ref[Vh:]) < 1e-6         # both solve to 1e-8: solutions agree to ~cond * 1e-8
    ctx.close()


@pytest.mark.parametrize("solver", ["CG", "BICGSTAB"])
def test_mdagm_solver(oracle, solver):
    """HMC-side shells: M^dag M psi = chi by CG (syssolver_mdagm_cg.h:59-94) and two-step BiCGStab
    (syssolver_mdagm_bicgstab.h:62-110); iteration counts against the CPU restatement, true residual of the normal system."""
    latt = (8, 8, 8, 8)
    u, op, ctx, cp = setup(oracle, latt, "double", gauge="weak")
    chi = fields.gaussian_fermion(latt, seed=14, cb=1)
    Vh = ctx.Vh
    rsd = 1e-8
    if solver == "CG":
        psi_ref, n_ref, resid_ref = op.solve_mdagm_cg(chi, np.zeros_like(chi), rsd, 2000)
        code = L.B200_SOLVER_CG
    else:
        psi_ref, n_ref, resid_ref = op.solve_mdagm_bicgstab(chi, np.zeros_like(chi), rsd, 2000)
        code = L.B200_SOLVER_BICGSTAB
    psi, info = ctx.invert_mdagm(chi[Vh:], None, solver=code, rsd=rsd, max_iter=2000)
    assert info.converged == 1
    assert abs(info.n_count - n_ref) <= max(2, 0.05 * n_ref), (info.n_count, n_ref)
    full = np.zeros_like(chi)
    full[Vh:] = psi
    r = chi - op.apply(op.apply(full, +1), -1)
    res = np.sqrt(np.sum(r[Vh:] ** 2))
    assert res / np.sqrt(np.sum(chi[Vh:] ** 2)) < 50 * rsd
    assert abs(info.resid - res) < 1e-3 * res + 1e-13
    assert rel_site_err(psi, psi_ref[Vh:]) < 1e-5
    ctx.close()


@pytest.mark.parametrize("mdagm", [False, True])
@pytest.mark.parametrize("delta", [0.1, 0.01])
def test_reliable_cg_mixed_precision(oracle, delta, mdagm):
    """RelInvCG_a (reliable_cg.cc:10-190) with fp32 inner / fp64 outer: reaches the fp64 target residual that a pure
    fp32 solve cannot, with an iteration count within a few percent (+2) of the CPU restatement of the same algorithm
    and of plain fp64 CG; at least one residual replacement must have happened."""
    latt = (8, 8, 8, 8)
    u, op, ctx, cp = setup(oracle, latt, "double", gauge="weak")
    chi = fields.gaussian_fermion(latt, seed=12, cb=1)
    Vh = ctx.Vh
    rsd = 1e-10
    psi_ref, n_ref, nupd_ref, resid_ref = op.solve_reliable_cg(chi, np.zeros_like(chi), rsd, delta, 2000, mdagm=mdagm)
    psi, info = ctx.invert_reliable(chi[Vh:], None, rsd=rsd, delta=delta, max_iter=2000, mdagm=mdagm)
    assert info.converged == 1
    assert abs(info.n_count - n_ref) <= max(3, 0.08 * n_ref), (info.n_count, n_ref)
    if mdagm:
        _, n64, _ = op.solve_mdagm_cg(chi, np.zeros_like(chi), rsd, 2000)
    else:
        _, n64, _, _ = op.solve_cg(chi, np.zeros_like(chi), rsd, 2000)
    assert info.n_count <= 1.15 * n64 + 3, (info.n_count, n64)
    assert info.n_updates >= 1 and abs(info.n_updates - nupd_ref) <= 2, (info.n_updates, nupd_ref)
    full = np.zeros_like(chi)
    full[Vh:] = psi
    r = chi - (op.apply(op.apply(full, +1), -1) if mdagm else op.apply(full, +1))
    rel = np.sqrt(np.sum(r[Vh:] ** 2) / np.sum(chi[Vh:] ** 2))
    assert rel < 20 * rsd, rel                                  # far below what fp32 alone can reach (~1e-6)
    assert abs(info.rel_resid - rel) < 1e-2 * rel + 1e-14
    assert rel_site_err(psi, psi_ref[Vh:]) < 1e-6
    # the second call reuses the fp32 twin; a float context must refuse
    psi2, info2 = ctx.invert_reliable(chi[Vh:], None, rsd=rsd, delta=delta, max_iter=2000, mdagm=mdagm)
    assert info2.n_count == info.n_count and np.array_equal(psi2, psi)
    ctx.close()
    u, op, ctxf, cp = setup(oracle, latt, "single", gauge="weak")
    with pytest.raises(L.B200Error) as e:
        ctxf.invert_reliable(chi[Vh:].astype(np.float32), None, rsd=1e-6, delta=0.1, max_iter=10)
    assert e.value.code == L.B200_ERR_ARG
    ctxf.close()


@pytest.mark.parametrize("mdagm", [False, True])
@pytest.mark.parametrize("delta", [0.1, 0.01])
def test_reliable_bicgstab_mixed_precision(oracle, delta, mdagm):
    """RelInvBiCGStab_a (reliable_bicgstab.cc:13-290), fp32 recurrences / fp64 residual replacement: reaches an fp64
    target residual fp32 alone cannot, iteration and replacement counts close to the CPU restatement of the same
    algorithm (fp32 BiCGStab is more erratic than CG, hence the wider band), two-step M^dag M variant included."""
    latt = (8, 8, 8, 8)
    u, op, ctx, cp = setup(oracle, latt, "double", gauge="weak")
    chi = fields.gaussian_fermion(latt, seed=12, cb=1)
    Vh = ctx.Vh
    rsd = 1e-10
    z = np.zeros_like(chi)
    if mdagm:
        psi_ref, n_ref, nupd_ref, _ = op.solve_mdagm_reliable_bicgstab(chi, z, rsd, delta, 2000)
        _, n64, _ = op.solve_mdagm_bicgstab(chi, z, rsd, 2000)
    else:
        psi_ref, n_ref, nupd_ref, _ = op.solve_reliable_bicgstab(chi, z, rsd, delta, 2000)
        _, n64, _, _ = op.solve_bicgstab(chi, z, rsd, 2000)
    psi, info = ctx.invert_reliable_bicgstab(chi[Vh:], None, rsd=rsd, delta=delta, max_iter=2000, mdagm=mdagm)
    assert info.converged == 1
    assert abs(info.n_count - n_ref) <= max(4, 0.15 * n_ref), (info.n_count, n_ref)
    assert info.n_count <= 1.25 * n64 + 4, (info.n_count, n64)
    assert info.n_updates >= 1 and abs(info.n_updates - nupd_ref) <= max(3, 0.3 * nupd_ref), (info.n_updates, nupd_ref)
    full = np.zeros_like(chi)
    full[Vh:] = psi
    r = chi - (op.apply(op.apply(full, +1), -1) if mdagm else op.apply(full, +1))
    rel = np.sqrt(np.sum(r[Vh:] ** 2) / np.sum(chi[Vh:] ** 2))
    assert rel < (200 if mdagm else 20) * rsd, rel              # far below what fp32 alone can reach (~1e-6)
    assert abs(info.rel_resid - rel) < 1e-2 * rel + 1e-14
    assert rel_site_err(psi, psi_ref[Vh:]) < 1e-6
    # deterministic; an exact initial guess returns at once; a float context must refuse
    psi2, info2 = ctx.invert_reliable_bicgstab(chi[Vh:], None, rsd=rsd, delta=delta, max_iter=2000, mdagm=mdagm)
    assert info2.n_count == info.n_count and np.array_equal(psi2, psi)
    psi3, info3 = ctx.invert_reliable_bicgstab(chi[Vh:], psi, rsd=1e-8, delta=delta, max_iter=2000, mdagm=mdagm)
    assert info3.n_count == 0 and info3.converged == 1 and np.array_equal(psi3, psi)
    ctx.close()
    u, op, ctxf, cp = setup(oracle, latt, "single", gauge="weak")
    with pytest.raises(L.B200Error) as e:
        ctxf.invert_reliable_bicgstab(chi[Vh:].astype(np.float32), None, rsd=1e-6, delta=0.1, max_iter=10)
    assert e.value.code == L.B200_ERR_ARG
    ctxf.close()


def test_plugin_mirror_drop_in(oracle):
    """LinOpSysSolverB200Clover used the way quarkprop4_w.cc:86-109 uses a LinOpSystemSolver: XML-like params in,
    (psi, chi) -> {n_count, resid}; GPU-built clover; non-convergence raises unless SilentFail (.h:634-644)."""
    latt = (8, 8, 8, 8)
    u = fields.apply_bc(latt, fields.weak_gauge(latt, seed=11))
    op = oracle.Op(latt, u, 0.1, 1.0)
    chi = fields.gaussian_fermion(latt, seed=12, cb=1)
    Vh = chi.shape[0] // 2
    p = SysSolverB200CloverParams(CloverParams=CloverFermActParams(Mass=0.1, clovCoeffR=1.0, clovCoeffT=1.0),
                                  RsdTarget=1e-8, MaxIter=1000, SolverType="BICGSTAB", AntiPeriodicT=True)
    S = LinOpSysSolverB200Clover(latt, u, p)
    psi = np.zeros((Vh, 4, 3, 2))
    res = S(psi, chi[Vh:])
    full = np.zeros_like(chi)
    full[Vh:] = psi
    r = chi - op.apply(full, +1)
    assert np.sqrt(np.sum(r[Vh:] ** 2)) == pytest.approx(res.resid, rel=1e-3)
    assert res.resid / np.sqrt(np.sum(chi[Vh:] ** 2)) < 1e-7 and res.n_count > 0
    S.close()
    p2 = SysSolverB200CloverParams(CloverParams=p.CloverParams, RsdTarget=1e-10, MaxIter=3, SolverType="CG")
    S2 = LinOpSysSolverB200Clover(latt, u, p2)
    with pytest.raises(SolverFailure):
        S2(np.zeros((Vh, 4, 3, 2)), chi[Vh:])
    p2.SilentFail = True
    assert S2(np.zeros((Vh, 4, 3, 2)), chi[Vh:]).n_count == 3
    S2.close()


def test_qprop_full_lattice(oracle):
    """12 spin-colour point sources through b200_qprop == the unpreconditioned system solved (eoprec_fermact_qprop.cc:41-80)."""
    latt = (4, 4, 4, 8)
    u, op, ctx, _ = setup(oracle, latt, "double", gauge="weak")
    srcs = np.stack([fields.point_source(latt, s, c) for s in range(4) for c in range(3)])
    sol, infos = ctx.qprop(srcs, solver=L.B200_SOLVER_CG, rsd=1e-10, max_iter=500)
    for i in range(12):
        r = op.unprec_apply(sol[i], +1) - srcs[i]
        assert np.linalg.norm(r) / np.linalg.norm(srcs[i]) < 1e-8
        assert infos[i].converged == 1
    ctx.close()


@pytest.mark.parametrize("prec", ["double", "single"])
@pytest.mark.parametrize("nrhs,l2_kb", [(12, None), (5, None), (12, 250), (7, 40)])
def test_multi_rhs_operator_parity(oracle, nrhs, l2_kb, prec, monkeypatch):
    """Batched fields through the multi-RHS kernels (one CTA = 32 sites x a group of right-hand sides, links shared through
    L1): every right-hand side of Dslash, A^-1, M and M^dagger equals the oracle applied to that source alone.  l2_kb
    shrinks the L2 budget so that the z-chunked traversal order (used on big lattices) is exercised on a small one."""
    if l2_kb is not None:
        monkeypatch.setenv("B200_MRHS_L2_KB", str(l2_kb))
    latt = (8, 4, 6, 6)        # Vh = 576 = 18 x 32; exercises wrap-around in every direction
    u, op, ctx, cp = setup(oracle, latt, prec, gauge="random")
    Vh = ctx.Vh
    tol = 1e-13 if prec == "double" else 2e-6
    srcs = np.stack([fields.gaussian_fermion(latt, seed=100 + i) for i in range(nrhs)])
    for out_cb in (0, 1):
        fin = ctx.mfield(nrhs, srcs[:, (1 - out_cb) * Vh:(2 - out_cb) * Vh].astype(NP[prec]))
        fout = ctx.mfield(nrhs)
        for isign in (+1, -1):
            ctx.dev_dslash(fout, fin, isign, out_cb)
            got = fout.download()
            for i in range(nrhs):
                want = op.dslash(srcs[i], isign, out_cb)[out_cb * Vh:(out_cb + 1) * Vh]
                assert rel_site_err(got[i], want) < tol
    fin = ctx.mfield(nrhs, srcs[:, Vh:].astype(NP[prec]))
    fout = ctx.mfield(nrhs)
    for isign in (+1, -1):
        ctx.dev_matpc(fout, fin, isign)
        got = fout.download()
        for i in range(nrhs):
            odd = srcs[i].copy()
            odd[:Vh] = 0
            assert rel_site_err(got[i], op.apply(odd, isign)[Vh:]) < 2 * tol
    n2 = ctx.dev_norm2(fin)
    for i in range(nrhs):
        want = np.sum(srcs[i, Vh:].astype(NP[prec]).astype(np.float64) ** 2)
        assert abs(n2[i] - want) < 1e-12 * want
    ctx.close()


@pytest.mark.parametrize("split", [False, True], ids=["ticket", "split"])
@pytest.mark.parametrize("solver", ["CG", "BICGSTAB"])
def test_multi_rhs_solvers_lockstep(oracle, solver, split, monkeypatch):
    """12 different right-hand sides solved in lockstep converge independently: each needs the iteration count of its own
    single-RHS solve (+-1: the batched reductions sum in a different, still fixed, order) and reaches the target residual.
    Includes a zero source (converged before the first iteration) and sources of very different norms.
    split: the reductions of the batched kernels finished by dslash_mrhs_finish_kernel (one CTA per right-hand side), the path
    big lattices take; B200_SPLIT_MIN_BLOCKS=0 forces it here."""
    latt = (8, 8, 8, 8)
    if split:
        monkeypatch.setenv("B200_SPLIT_MIN_BLOCKS", "0")
    u, op, ctx, cp = setup(oracle, latt, "double", gauge="weak")
    Vh = ctx.Vh
    code = L.B200_SOLVER_CG if solver == "CG" else L.B200_SOLVER_BICGSTAB
    rsd = 1e-9
    srcs = []
    for i in range(12):
        f = fields.gaussian_fermion(latt, seed=200 + i, cb=1)[Vh:] if i % 3 else fields.point_source(latt, i % 4, i % 3)[:Vh].copy()
        srcs.append(f * 10.0 ** (i - 6))
    if solver == "CG":
        srcs[7] = np.zeros_like(srcs[7])
    srcs = np.stack(srcs)
    chi = ctx.mfield(12, srcs)
    psi = ctx.mfield(12)
    infos = ctx.dev_invert(psi, chi, solver=code, rsd=rsd, max_iter=2000)
    sol = psi.download()
    for i in range(12):
        if not srcs[i].any():
            assert infos[i].converged == 1 and infos[i].n_count == 0 and not sol[i].any()
            continue
        one, info1 = ctx.invert(srcs[i], None, solver=code, rsd=rsd, max_iter=2000)
        assert infos[i].converged == 1
        assert abs(infos[i].n_count - info1.n_count) <= 1, (i, infos[i].n_count, info1.n_count)
        full = np.zeros((2 * Vh, 4, 3, 2))
        full[Vh:] = sol[i]
        rhs = np.zeros_like(full)
        rhs[Vh:] = srcs[i]
        r = rhs - op.apply(full, +1)
        rel = np.sqrt(np.sum(r[Vh:] ** 2) / np.sum(srcs[i] ** 2))
        assert rel < 20 * rsd, (i, rel)
        assert abs(infos[i].rel_resid - rel) < 1e-2 * rel + 1e-14
        assert rel_site_err(sol[i], one) < 1e-6
    ctx.close()


def test_qprop_batches(oracle, monkeypatch):
    """b200_qprop in batches that do not divide the number of sources (12 = 5 + 5 + 2) gives the same propagator."""
    latt = (4, 4, 4, 8)
    u, op, ctx, _ = setup(oracle, latt, "double", gauge="weak")
    srcs = np.stack([fields.point_source(latt, s, c) for s in range(4) for c in range(3)])
    sol12, _ = ctx.qprop(srcs, solver=L.B200_SOLVER_BICGSTAB, rsd=1e-10, max_iter=500)
    monkeypatch.setenv("B200_QPROP_BATCH", "5")
    sol5, infos = ctx.qprop(srcs, solver=L.B200_SOLVER_BICGSTAB, rsd=1e-10, max_iter=500)
    monkeypatch.setenv("B200_QPROP_BATCH", "1")
    sol1, _ = ctx.qprop(srcs, solver=L.B200_SOLVER_BICGSTAB, rsd=1e-10, max_iter=500)
    for i in range(12):
        assert infos[i].converged == 1
        r = op.unprec_apply(sol5[i], +1) - srcs[i]
        assert np.linalg.norm(r) / np.linalg.norm(srcs[i]) < 1e-8
        assert np.abs(sol5[i] - sol12[i]).max() < 1e-9 and np.abs(sol1[i] - sol12[i]).max() < 1e-9
    ctx.close()


# ---------------------------------------------------------------------------------------- SURVEY section 8 (f4)
@pytest.mark.parametrize("prec", ["double", "single"])
@pytest.mark.parametrize("recon,aniso,gpu_clover", [(L.B200_RECONS_NONE, False, False), (L.B200_RECONS_12, True, True)])
def test_symmetric_operator_parity(oracle, prec, recon, aniso, gpu_clover):
    """SymEvenOddPrecCloverLinOp::operator() (seoprec_clover_linop_w.cc:147-193) PLUS and MINUS against the restatement,
    A_oo^-1 built on the device (invclov.choles(1)) against the restated ldagdlinv, log det A_oo, and the switch back."""
    latt = (6, 4, 4, 8)
    u, op, ctx, cp = setup(oracle, latt, prec, recon=recon, aniso=aniso, gpu_clover=gpu_clover)
    Vh = ctx.Vh
    psi = fields.gaussian_fermion(latt, seed=21, cb=1)
    x = psi[Vh:].astype(NP[prec])
    asym_plus = ctx.matpc(x, +1)
    ctx.set_preconditioning(True)
    op.set_symmetric(True)
    tol = TOL[prec] * (10 if recon == L.B200_RECONS_12 else 1)
    for isign in (+1, -1):
        got = ctx.matpc(x, isign)
        want = op.apply(psi, isign)[Vh:]
        assert rel_site_err(got, want) < tol, (isign, rel_site_err(got, want))
    g = oracle.Geom(latt)
    got = ctx.clover_apply(x, 1, inverse=True)
    want = oracle.clover_apply(g, psi, op.invclov, 1)[Vh:]
    assert rel_site_err(got, want) < (1e-12 if prec == "double" else 2e-6)
    if prec == "double":
        assert ctx.clover_logdet(1) == pytest.approx(op.tr_log(1), rel=1e-12)
    # device-resident path + gamma5-hermiticity <chi, S psi> = <S^dag chi, psi>
    chi = fields.gaussian_fermion(latt, seed=22, cb=1)
    fx, fc, fo, fo2 = ctx.field(x), ctx.field(chi[Vh:].astype(NP[prec])), ctx.field(), ctx.field()
    ctx.dev_matpc(fo, fx, +1)
    ctx.dev_matpc(fo2, fc, -1)
    a, b = ctx.dev_inner(fc, fo), ctx.dev_inner(fo2, fx)
    assert abs(a - b) < (1e-11 if prec == "double" else 1e-4) * abs(a)
    # batched fields run the multi-RHS kernels with the symmetric epilogues (MODE_SYM_PLUS / MODE_SYM_MINUS + the A_oo^-1 pass)
    nr = 5
    srcs = np.stack([fields.gaussian_fermion(latt, seed=60 + i, cb=1) for i in range(nr)])
    fin, fout = ctx.mfield(nr, srcs[:, Vh:].astype(NP[prec])), ctx.mfield(nr)
    for isign in (+1, -1):
        ctx.dev_matpc(fout, fin, isign)
        outs = fout.download()
        for i in range(nr):
            assert rel_site_err(outs[i], op.apply(srcs[i], isign)[Vh:]) < 2 * tol, (isign, i)
    infos = ctx.dev_invert(fout, fin, solver=L.B200_SOLVER_BICGSTAB, rsd=1e-8 if prec == "double" else 1e-5, max_iter=500)
    sols = fout.download()
    for i in range(nr):
        full = np.zeros_like(srcs[i]); full[Vh:] = sols[i]
        r = srcs[i] - op.apply(full, +1)
        assert infos[i].converged == 1 and np.sqrt(np.sum(r[Vh:] ** 2) / np.sum(srcs[i][Vh:] ** 2)) < (2e-7 if prec == "double" else 2e-4)
    # the asymmetric operator comes back bit-identical
    ctx.set_preconditioning(False)
    assert np.array_equal(ctx.matpc(x, +1), asym_plus)
    with pytest.raises(L.B200Error):
        ctx.clover_apply(x, 1, inverse=True)
    ctx.close()


@pytest.mark.parametrize("solver", ["CG", "BICGSTAB", "MDAGM_CG", "RELIABLE", "RELIABLE_BICGSTAB"])
def test_symmetric_solvers(oracle, solver):
    """Every solver shell on the symmetric operator: iteration counts against the restated loops run on the restated
    SymEvenOddPrecCloverLinOp, true residual recomputed on the CPU (symm_prec_tests.cc:210-249 is the reference's check)."""
    latt = (8, 8, 8, 8)
    u, op, ctx, cp = setup(oracle, latt, "double", gauge="weak")
    ctx.set_preconditioning(True)
    op.set_symmetric(True)
    chi = fields.gaussian_fermion(latt, seed=12, cb=1)
    Vh = ctx.Vh
    z = np.zeros_like(chi)
    rsd = 1e-8
    mdagm = solver == "MDAGM_CG"
    if solver == "CG":
        ref, n_ref, _, _ = op.solve_cg(chi, z, rsd, 2000)
        psi, info = ctx.invert(chi[Vh:], None, solver=L.B200_SOLVER_CG, rsd=rsd, max_iter=2000)
    elif solver == "BICGSTAB":
        ref, n_ref, _, _ = op.solve_bicgstab(chi, z, rsd, 2000)
        psi, info = ctx.invert(chi[Vh:], None, solver=L.B200_SOLVER_BICGSTAB, rsd=rsd, max_iter=2000)
    elif solver == "MDAGM_CG":
        ref, n_ref, _ = op.solve_mdagm_cg(chi, z, rsd, 2000)
        psi, info = ctx.invert_mdagm(chi[Vh:], None, solver=L.B200_SOLVER_CG, rsd=rsd, max_iter=2000)
    elif solver == "RELIABLE":
        ref, n_ref, _, _ = op.solve_reliable_cg(chi, z, rsd, 0.1, 2000)
        psi, info = ctx.invert_reliable(chi[Vh:], None, rsd=rsd, delta=0.1, max_iter=2000)
    else:
        ref, n_ref, _, _ = op.solve_reliable_bicgstab(chi, z, rsd, 0.1, 2000)
        psi, info = ctx.invert_reliable_bicgstab(chi[Vh:], None, rsd=rsd, delta=0.1, max_iter=2000)
    assert info.converged == 1
    assert abs(info.n_count - n_ref) <= max(4 if "BICGSTAB" in solver else 3, 0.08 * n_ref), (info.n_count, n_ref)
    full = np.zeros_like(chi)
    full[Vh:] = psi
    r = chi - (op.apply(op.apply(full, +1), -1) if mdagm else op.apply(full, +1))
    rel = np.sqrt(np.sum(r[Vh:] ** 2) / np.sum(chi[Vh:] ** 2))
    assert rel < 50 * rsd
    assert abs(info.rel_resid - rel) < 1e-2 * rel + 1e-14
    assert rel_site_err(psi, ref[Vh:]) < 1e-5
    ctx.close()


def test_symmetric_qprop_and_plugin(oracle, monkeypatch):
    """SymEvenOddPrecActQprop (seoprec_fermact_qprop.cc:41-100) on the device solves the same unpreconditioned system;
    the plugin mirror with SymmetricLinop solves the caller's symmetric A."""
    latt = (4, 4, 4, 8)
    u, op, ctx, _ = setup(oracle, latt, "double", gauge="weak")
    ctx.set_preconditioning(True)
    srcs = np.stack([fields.point_source(latt, s, c) for s, c in ((0, 0), (3, 2))] + [fields.gaussian_fermion(latt, seed=31)])
    launches = {}
    for batch in ("12", "1"):          # all right-hand sides through the batched symmetric kernels, then one at a time
        monkeypatch.setenv("B200_QPROP_BATCH", batch)
        l0 = ctx.launch_count
        sol, infos = ctx.qprop(srcs, solver=L.B200_SOLVER_BICGSTAB, rsd=1e-10, max_iter=500)
        launches[batch] = ctx.launch_count - l0
        for i in range(len(srcs)):
            r = op.unprec_apply(sol[i], +1) - srcs[i]
            assert np.linalg.norm(r) / np.linalg.norm(srcs[i]) < 1e-8
            assert infos[i].converged == 1
    assert launches["12"] < 0.6 * launches["1"]      # the batch really went through in lockstep
    ctx.close()
    op.set_symmetric(True)
    chi = fields.gaussian_fermion(latt, seed=32, cb=1)
    Vh = chi.shape[0] // 2
    p = SysSolverB200CloverParams(CloverParams=CloverFermActParams(Mass=0.1, clovCoeffR=1.0, clovCoeffT=1.0),
                                  RsdTarget=1e-9, MaxIter=1000, SolverType="CG", SymmetricLinop=True)
    S = LinOpSysSolverB200Clover(latt, u, p)
    psi = np.zeros((Vh, 4, 3, 2))
    res = S(psi, chi[Vh:])
    full = np.zeros_like(chi)
    full[Vh:] = psi
    r = chi - op.apply(full, +1)
    assert np.sqrt(np.sum(r[Vh:] ** 2)) == pytest.approx(res.resid, rel=1e-3)
    assert res.resid / np.sqrt(np.sum(chi[Vh:] ** 2)) < 1e-8
    S.close()


@pytest.mark.parametrize("prec,symmetric", [("double", False), ("double", True), ("single", False)])
def test_multishift_cg(oracle, prec, symmetric):
    """MInvCG2_a (minvcg2.cc:74-373): same iteration count as the CPU restatement, every shift's TRUE residual below its
    own target, solutions equal to the restatement's, shifts converge (and freeze) independently."""
    latt = (8, 8, 8, 8)
    u, op, ctx, cp = setup(oracle, latt, prec, gauge="weak")
    ctx.set_preconditioning(symmetric)
    op.set_symmetric(symmetric)
    chi = fields.gaussian_fermion(latt, seed=12, cb=1)
    Vh = ctx.Vh
    shifts = [0.0005, 0.01, 0.08, 0.6, 3.0]
    rsd = [1e-8, 1e-8, 1e-7, 1e-6, 1e-5] if prec == "double" else [1e-5, 1e-5, 1e-5, 1e-4, 1e-4]
    ref, n_ref, rel_ref = op.solve_multishift(chi, shifts, rsd, 2000)
    psi, infos = ctx.invert_multishift(chi[Vh:].astype(NP[prec]), shifts, rsd, max_iter=2000)
    assert all(i.converged == 1 for i in infos)
    assert len({i.n_count for i in infos}) == 1
    assert abs(infos[0].n_count - n_ref) <= max(2, 0.05 * n_ref), (infos[0].n_count, n_ref)
    for s, sh in enumerate(shifts):
        full = np.zeros_like(chi)
        full[Vh:] = psi[s]
        r = chi - op.apply(op.apply(full, +1), -1) - sh * full
        rel = np.sqrt(np.sum(r[Vh:] ** 2) / np.sum(chi[Vh:] ** 2))
        assert rel < (10 if prec == "double" else 100) * rsd[s], (s, rel)
        assert abs(infos[s].rel_resid - rel) < 2e-2 * rel + (1e-14 if prec == "double" else 1e-7)
        assert rel_site_err(psi[s], ref[s][Vh:]) < (1e-5 if prec == "double" else 5e-3)
    # one shift == ordinary CG on M^dag M + sigma; zero source returns zero in zero iterations (minvcg2.cc:135-148)
    one, inf1 = ctx.invert_multishift(chi[Vh:].astype(NP[prec]), [0.08], rsd[2], max_iter=2000)
    assert rel_site_err(one[0], psi[2]) < (1e-6 if prec == "double" else 5e-3) and inf1[0].n_count <= infos[0].n_count
    zero, inf0 = ctx.invert_multishift(np.zeros((Vh, 4, 3, 2), dtype=NP[prec]), shifts, rsd, max_iter=50)
    assert inf0[0].n_count == 0 and not zero.any()
    # not converging is reported, not fatal; too many shifts are refused
    _, infx = ctx.invert_multishift(chi[Vh:].astype(NP[prec]), shifts, 1e-12, max_iter=3)
    assert infx[0].converged == 0 and infx[0].n_count == 3
    with pytest.raises(L.B200Error) as e:
        ctx.invert_multishift(chi[Vh:].astype(NP[prec]), np.linspace(0.1, 1, L.B200_MAX_SHIFTS + 1), 1e-6, max_iter=3)
    assert e.value.code == L.B200_ERR_ARG
    ctx.close()


def test_multishift_plugin_mirror_and_device_fields(oracle):
    """MdagMMultiSysSolverCG::operator()(psi[], shifts, chi) mirror, and the device-resident entry with 16 shifts
    (more vectors than the batched operators take: storage-only batched field)."""
    latt = (4, 4, 4, 8)
    u = fields.apply_bc(latt, fields.weak_gauge(latt, seed=11))
    op = oracle.Op(latt, u, 0.1, 1.0)
    chi = fields.gaussian_fermion(latt, seed=12, cb=1)
    Vh = chi.shape[0] // 2
    p = SysSolverB200CloverParams(CloverParams=CloverFermActParams(Mass=0.1, clovCoeffR=1.0, clovCoeffT=1.0),
                                  RsdTarget=1e-9, MaxIter=1000)
    S = MdagMMultiSysSolverB200Clover(latt, u, p)
    shifts = list(np.geomspace(1e-3, 5.0, 16))
    psi, res = S(shifts, chi[Vh:])
    ref, n_ref, _ = op.solve_multishift(chi, shifts, 1e-9, 1000)
    assert abs(res.n_count - n_ref) <= 2
    for s in range(16):
        assert rel_site_err(psi[s], ref[s][Vh:]) < 1e-6
    ctx = S.ctx
    fpsi, fchi = ctx.mfield(16), ctx.field(chi[Vh:])
    infos = ctx.dev_invert_multishift(fpsi, fchi, shifts, 1e-9, max_iter=1000)
    assert infos[0].n_count == res.n_count
    assert np.array_equal(fpsi.download(irhs=7), psi[7])
    with pytest.raises(L.B200Error):                      # 16 vectors: storage only
        ctx.dev_matpc(ctx.mfield(16), fpsi, +1)
    p.MaxIter = 2
    with pytest.raises(SolverFailure):
        S(shifts, chi[Vh:])
    S.close()


def test_multishift_aniso_recon12_gpu_clover(oracle):
    """Multi-shift CG on an anisotropic operator with 12-real links and the GPU-built clover term."""
    latt = (6, 4, 4, 8)
    u, op, ctx, cp = setup(oracle, latt, "double", gauge="weak", recon=L.B200_RECONS_12, aniso=True, gpu_clover=True)
    chi = fields.gaussian_fermion(latt, seed=15, cb=1)
    Vh = ctx.Vh
    shifts = [0.002, 0.1, 1.0]
    ref, n_ref, _ = op.solve_multishift(chi, shifts, 1e-9, 2000)
    psi, infos = ctx.invert_multishift(chi[Vh:], shifts, 1e-9, max_iter=2000)
    assert all(i.converged == 1 for i in infos) and abs(infos[0].n_count - n_ref) <= max(2, 0.05 * n_ref)
    for s_ in range(3):
        assert infos[s_].rel_resid < 1e-8
        assert rel_site_err(psi[s_], ref[s_][Vh:]) < 1e-6
    ctx.close()


@pytest.mark.parametrize("threads,pin_kb", [("0", None), ("3", "64"), ("2", "36")])
def test_host_copy_paths_agree(oracle, threads, pin_kb, monkeypatch):
    """Pageable host buffers go through the pinned bounce pipeline (a thread team fills one pinned buffer while the DMA
    engine drains the other); pinned ones (b200_host_alloc) and B200_COPY_THREADS=0 take plain cudaMemcpy: same bits on
    the device either way, up and down.  B200_PIN_KB shrinks the bounce buffers so that gauge, clover and fermion
    transfers of this small lattice span many pieces (36 KB: pieces that are no multiple of a site record)."""
    import ctypes as C
    monkeypatch.setenv("B200_COPY_THREADS", threads)
    if pin_kb:
        monkeypatch.setenv("B200_PIN_KB", pin_kb)
    latt = (16, 8, 8, 12)
    u, op, ctx, cp = setup(oracle, latt, "double")
    Vh = ctx.Vh
    psi = fields.gaussian_fermion(latt, seed=23, cb=1)
    want = op.apply(psi, +1)[Vh:]
    # pageable in, pageable out
    got = ctx.matpc(psi[Vh:].copy(), +1)
    assert rel_site_err(got, want) < TOL["double"]
    # pinned in, pinned out
    nbytes = Vh * 24 * 8
    pin_in, pin_out = C.c_void_p(), C.c_void_p()
    L.check(ctx.lib.b200_host_alloc(C.byref(pin_in), nbytes))
    L.check(ctx.lib.b200_host_alloc(C.byref(pin_out), nbytes))
    a_in = np.ctypeslib.as_array((C.c_double * (Vh * 24)).from_address(pin_in.value)).reshape(Vh, 4, 3, 2)
    a_out = np.ctypeslib.as_array((C.c_double * (Vh * 24)).from_address(pin_out.value)).reshape(Vh, 4, 3, 2)
    a_in[...] = psi[Vh:]
    L.check(ctx.lib.b200_clover_matpc(ctx.h, pin_out, pin_in, L.B200_DOUBLE, +1))
    assert np.array_equal(a_out, got)
    del a_in, a_out
    ctx.lib.b200_host_free(pin_in)
    ctx.lib.b200_host_free(pin_out)
    ctx.close()


def test_error_behaviour():
    """Bad arguments fail loudly with a code and a message (no exceptions cross the C ABI)."""
    lib = L.load()
    with pytest.raises(L.B200Error) as e:
        Context((5, 4, 4, 4))                       # odd global extent (shift_table_scalar.cc:23-28)
    assert e.value.code == L.B200_ERR_ARG
    ctx = Context((4, 4, 4, 4))
    f = ctx.field()
    g2 = ctx.field()
    with pytest.raises(L.B200Error) as e:
        ctx.dev_dslash(f, g2, +1, 0)                # gauge not loaded
    assert e.value.code == L.B200_ERR_STATE
    u = fields.unit_gauge((4, 4, 4, 4))
    ctx.load_gauge(u)
    with pytest.raises(L.B200Error) as e:
        ctx.dev_matpc(f, g2, +1)                    # clover not loaded
    assert e.value.code == L.B200_ERR_STATE
    with pytest.raises(L.B200Error) as e:
        ctx.dev_dslash(f, g2, 0, 0)                 # isign must be +-1
    assert e.value.code == L.B200_ERR_ARG
    ctx.close()


@pytest.mark.parametrize("prec", ["double", "single"])
@pytest.mark.parametrize("symmetric", [False, True])
def test_twisted_mass_operator_parity(oracle, prec, symmetric):
    """a10: chi += (+/-) mu i gamma_5 psi behind both operators (eoprec_clover_linop_w.cc:174-184,
    seoprec_clover_linop_w.cc:174-184) -- single- and multi-RHS kernels against the restated operator, <chi,M psi> =
    <M^dag chi,psi>, and mu = 0 restores the untwisted operator bit for bit."""
    latt = (8, 4, 4, 8)
    u, op, ctx, _ = setup(oracle, latt, prec, gauge="weak")
    Vh = ctx.Vh
    psi = fields.gaussian_fermion(latt, seed=21, cb=1)
    chi = fields.gaussian_fermion(latt, seed=22, cb=1)
    if symmetric:
        ctx.set_preconditioning(True)
        op.set_symmetric(True)
    plain = {s: ctx.matpc(psi[Vh:].astype(NP[prec]), s) for s in (+1, -1)}
    mu = 0.137
    ctx.set_twisted_mass(mu)
    op.set_twisted_mass(mu)
    got = {}
    for isign in (+1, -1):
        want = op.apply(psi, isign)[Vh:]
        got[isign] = ctx.matpc(psi[Vh:].astype(NP[prec]), isign)
        assert rel_site_err(got[isign], want) < 2 * TOL[prec], isign
        assert rel_site_err(got[isign], plain[isign]) > 1e-3          # the term is really there
    # <chi, M psi> = <M^dag chi, psi>
    mdag_chi = ctx.matpc(chi[Vh:].astype(NP[prec]), -1).astype(np.float64)
    c = lambda a: a[..., 0] + 1j * a[..., 1]
    lhs = np.vdot(c(chi[Vh:]), c(got[+1].astype(np.float64)))
    rhs = np.vdot(c(mdag_chi), c(psi[Vh:]))
    assert abs(lhs - rhs) < (1e-11 if prec == "double" else 2e-4) * abs(lhs)
    # batched kernels carry the term as well (asymmetric operator only)
    if not symmetric:
        nr = 3
        srcs = np.stack([fields.gaussian_fermion(latt, seed=30 + i, cb=1) for i in range(nr)])
        fin, fout = ctx.mfield(nr, srcs[:, Vh:].astype(NP[prec])), ctx.mfield(nr)
        ctx.dev_matpc(fout, fin, +1)
        outs = fout.download()
        for i in range(nr):
            assert rel_site_err(outs[i], op.apply(srcs[i], +1)[Vh:]) < 2 * TOL[prec]
    # solvers run on the twisted operator: true residual with the restated one
    sol, info = ctx.invert(chi[Vh:].astype(NP[prec]), None, solver=L.B200_SOLVER_BICGSTAB, rsd=1e-8 if prec == "double" else 1e-5, max_iter=500)
    full = np.zeros_like(chi)
    full[Vh:] = sol
    res = chi - op.apply(full, +1)
    assert info.converged and np.sqrt(np.sum(res[Vh:] ** 2) / np.sum(chi[Vh:] ** 2)) < (2e-7 if prec == "double" else 2e-4)
    ctx.set_twisted_mass(0.0)
    for isign in (+1, -1):
        assert np.array_equal(ctx.matpc(psi[Vh:].astype(NP[prec]), isign), plain[isign])
    ctx.close()


@pytest.mark.parametrize("solver", ["CG", "BICGSTAB"])
def test_split_reduction_path(oracle, solver, monkeypatch):
    """Large single-RHS grids let the CTAs store their partial sums and a one-CTA kernel finish the reduction
    (ReduceBuf::split, reduce.cuh); B200_SPLIT_MIN_BLOCKS=0 forces that path on a small lattice.  Same iteration count as
    the CPU restatement, bit-identical solution to the in-kernel ticket path (both sum in the same two-level order only
    above RED_FLAT_MAX blocks, so here: equal to rounding), true residual below target."""
    latt = (8, 8, 8, 8)
    code = L.B200_SOLVER_CG if solver == "CG" else L.B200_SOLVER_BICGSTAB
    u, op, ctx, _ = setup(oracle, latt, "double", gauge="weak")
    Vh = ctx.Vh
    chi = fields.gaussian_fermion(latt, seed=41, cb=1)
    ref_sol, ref_info = ctx.invert(chi[Vh:], None, solver=code, rsd=1e-9, max_iter=500)
    ctx.close()
    monkeypatch.setenv("B200_SPLIT_MIN_BLOCKS", "0")
    u, op, ctx, _ = setup(oracle, latt, "double", gauge="weak")
    sol, info = ctx.invert(chi[Vh:], None, solver=code, rsd=1e-9, max_iter=500)
    n_launch = ctx.launch_count
    ctx.close()
    assert info.converged and abs(info.n_count - ref_info.n_count) <= 1
    assert np.abs(sol - ref_sol).max() < 1e-9 * np.abs(ref_sol).max()
    full = np.zeros_like(chi); full[Vh:] = sol
    res = chi - op.apply(full, +1)
    assert np.sqrt(np.sum(res[Vh:] ** 2) / np.sum(chi[Vh:] ** 2)) < 2e-8


def test_zero_initial_guess_is_not_copied(oracle):
    """b200_invert sets an all-zero initial guess on the device instead of copying it (host_all_zero, engine_impl.cuh; the
    propagator call site passes psi = zero, quarkprop4_w.cc:74).  The result must be bit-identical to the copied path,
    which a single -0.0 in the guess forces (its bit pattern is not zero), and a real guess must still be honoured."""
    latt = (8, 8, 8, 8)
    u, op, ctx, _ = setup(oracle, latt, "double", gauge="weak")
    Vh = ctx.Vh
    chi = fields.gaussian_fermion(latt, seed=51, cb=1)[Vh:]
    zero = np.zeros_like(chi)
    negzero = zero.copy(); negzero[Vh // 2, 1, 2, 0] = -0.0
    assert negzero.tobytes() != zero.tobytes()
    s0, i0 = ctx.invert(chi, zero, solver=L.B200_SOLVER_CG, rsd=1e-9, max_iter=500)
    s1, i1 = ctx.invert(chi, negzero, solver=L.B200_SOLVER_CG, rsd=1e-9, max_iter=500)
    assert i0.n_count == i1.n_count and np.array_equal(s0, s1)
    # starting from the solution: converged at once (the guess was used, not replaced by zero)
    s2, i2 = ctx.invert(chi, s0, solver=L.B200_SOLVER_CG, rsd=1e-8, max_iter=500)
    assert i2.n_count <= 1 and i2.converged
    ctx.close()


def test_recon12_rejects_links_it_cannot_rebuild(oracle):
    """12-real compression keeps rows 0,1 of every link and rebuilds row 2 as conj(row0 x row1): right only for SU(3) links
    whose one extra phase is the antiperiodic T boundary the engine carries itself.  The upload checks every link
    (recon12_check_kernel) and refuses anything else with B200_ERR_ARG instead of producing diag(1,1,-1) U silently:
    spatially antiperiodic fermion boundaries, a wrong t_boundary, non-unitary links."""
    latt = (4, 4, 4, 8)
    u0 = fields.random_gauge(latt, seed=71)
    ctx = Context(latt, prec="double")
    ctx.load_gauge(fields.apply_bc(latt, u0, (1, 1, 1, -1)), t_boundary=-1, reconstruct=L.B200_RECONS_12)      # fine
    ctx.load_gauge(fields.apply_bc(latt, u0, (1, 1, 1, 1)), t_boundary=+1, reconstruct=L.B200_RECONS_12)       # fine
    for bc, tb in (((-1, 1, 1, -1), -1),        # antiperiodic in x as well
                   ((1, 1, 1, -1), +1),         # the links carry the T phase but the caller says periodic
                   ((1, 1, 1, 1), -1)):         # ... and the other way round
        with pytest.raises(L.B200Error) as e:
            ctx.load_gauge(fields.apply_bc(latt, u0, bc), t_boundary=tb, reconstruct=L.B200_RECONS_12)
        assert e.value.code == L.B200_ERR_ARG and "RECONS_12" in str(e.value)
    bad = u0.copy(); bad[2, 17] *= 1.01                                  # one non-unitary link
    with pytest.raises(L.B200Error):
        ctx.load_gauge(bad, t_boundary=+1, reconstruct=L.B200_RECONS_12)
    ctx.load_gauge(fields.apply_bc(latt, u0, (-1, 1, 1, -1)), t_boundary=-1, reconstruct=L.B200_RECONS_NONE)   # uncompressed links take any phase
    ctx.close()
