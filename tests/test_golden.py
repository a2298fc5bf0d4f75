"""Golden vectors produced by the reference's own Dslash<double/float> and CloverSchur4D<double> (tests/golden/*.npz,
written by tests/golden/make_golden.py in the build container) against (i) the oracle restatement -- CPU, runs anywhere --
and (ii) the CUDA kernels through the C ABI -- GPU.  Unlike tests/test_oracle.py these need neither /root/reference nor
oracle/_ref at run time.
"""
import glob
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "ref_*.npz")))
KEYS = [(+1, 0, "p_cb0"), (+1, 1, "p_cb1"), (-1, 0, "m_cb0"), (-1, 1, "m_cb1")]


def rel_site_err(a, b):
    a = np.asarray(a, dtype=np.float64).reshape(a.shape[0], -1)
    b = np.asarray(b, dtype=np.float64).reshape(b.shape[0], -1)
    nb = np.linalg.norm(b, axis=1)
    m = nb > 0
    return float((np.linalg.norm(a - b, axis=1)[m] / nb[m]).max())


def test_fixtures_present():
    assert len(FILES) >= 2, "tests/golden/ref_*.npz missing"
    assert any("clov" in np.load(f).files for f in FILES)


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_oracle_dslash_matches_golden(oracle, path):
    z = np.load(path)
    L = tuple(int(v) for v in z["L"])
    g = oracle.Geom(L)
    pk = oracle.pack_gauge(L, z["u"], (1.0, 1.0, 1.0, 1.0))
    for isign, cb, key in KEYS:
        tgt = slice((1 - cb) * g.Vh, (2 - cb) * g.Vh)
        mine = oracle.dslash(g, z["psi"], pk, isign, cb)[tgt]
        assert rel_site_err(mine, z["dslash_d_" + key]) < 1e-14, key            # fp64 reference
        assert rel_site_err(mine, z["dslash_f_" + key]) < 1e-6, key             # fp32 reference (north_star tolerance)


@pytest.mark.parametrize("path", [f for f in FILES if "clov" in np.load(f).files], ids=os.path.basename)
def test_oracle_schur_matches_golden(oracle, path):
    """Restated EvenOddPrecCloverLinOp (eoprec_clover_linop_w.cc:142-187) == reference CloverSchur4D output."""
    z = np.load(path)
    L = tuple(int(v) for v in z["L"])
    g = oracle.Geom(L)
    op = oracle.Op(L, z["u"], 0.1, 1.0)
    # the fixture's clover term (with the one defective entry zeroed) must be what the restated build produces
    clov = op.clov.copy(); clov[:, 41] = 0.0
    assert np.abs(clov - z["clov"]).max() < 1e-13
    op.clov[:, 41] = 0.0
    op.invclov[:, 41] = 0.0
    assert np.abs(op.invclov[:g.Vh] - z["invclov_ee"]).max() < 1e-13
    chi = z["psi"].copy(); chi[:g.Vh] = 0.0
    for isign, key in ((+1, "p"), (-1, "m")):
        assert rel_site_err(op.apply(chi, isign)[g.Vh:], z["schur_d_" + key]) < 1e-13, key


# ------------------------------------------------------------------------------------------------ CUDA path
@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["double", "single"])
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_cuda_dslash_matches_golden(path, prec):
    from chroma_b200.solver import Context
    z = np.load(path)
    L = tuple(int(v) for v in z["L"])
    npdt = np.float64 if prec == "double" else np.float32
    ctx = Context(L, prec=prec)
    ctx.load_gauge(z["u"].astype(npdt), t_boundary=-1)
    Vh = ctx.Vh
    for isign, cb, key in KEYS:
        src = z["psi"][cb * Vh:(cb + 1) * Vh].astype(npdt)
        got = ctx.dslash(src, isign, 1 - cb)
        if prec == "double":
            assert rel_site_err(got, z["dslash_d_" + key]) < 1e-13, key
        else:
            assert rel_site_err(got, z["dslash_d_" + key]) < 1e-6, key
            assert rel_site_err(got, z["dslash_f_" + key]) < 2e-6, key          # both sides round in fp32
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["double", "single"])
@pytest.mark.parametrize("path", [f for f in FILES if "clov" in np.load(f).files], ids=os.path.basename)
def test_cuda_matpc_matches_golden(path, prec):
    from chroma_b200.solver import Context
    z = np.load(path)
    L = tuple(int(v) for v in z["L"])
    npdt = np.float64 if prec == "double" else np.float32
    ctx = Context(L, prec=prec)
    ctx.load_gauge(z["u"].astype(npdt), t_boundary=-1)
    Vh = ctx.Vh
    inv = np.zeros_like(z["clov"]); inv[:Vh] = z["invclov_ee"]
    ctx.load_clover(z["clov"].astype(npdt), inv.astype(npdt))
    for isign, key in ((+1, "p"), (-1, "m")):
        got = ctx.matpc(z["psi"][Vh:].astype(npdt), isign)
        assert rel_site_err(got, z["schur_d_" + key]) < (1e-13 if prec == "double" else 2e-6), key
    ctx.close()


# ------------------------------------------------------------------------------------------------ Chroma-level fixtures
# tests/golden/chroma_*.npz: outputs of the reference's mesField, clover site loops and solver loops (compiled unmodified,
# oracle/_ref/libref_chroma.so; written by make_golden.py::main_chroma).
CHROMA_FILES = sorted(glob.glob(os.path.join(HERE, "golden", "chroma_*.npz")))


def _chroma_case(z):
    L = tuple(int(v) for v in z["L"])
    an = dict(anisoP=True, t_dir=3, xi_0=float(z["xi_0"]), nu=float(z["nu"])) if int(z["aniso"]) else {}
    return L, an, float(z["clovCoeffR"]), float(z["clovCoeffT"])


def _close(a, b, tol):
    return np.abs(np.asarray(a, dtype=np.float64) - b).max() <= tol * np.abs(b).max()


@pytest.mark.parametrize("path", CHROMA_FILES, ids=os.path.basename)
def test_oracle_matches_chroma_golden(oracle, path):
    """The restated clover build and solver loops against the committed outputs of the reference's own code."""
    z = np.load(path)
    L, an, cR, cT = _chroma_case(z)
    g = oracle.Geom(L)
    Vh = g.Vh
    f = oracle.mesfield(g, z["u"])
    assert np.array_equal(f, z["f"])
    op = oracle.Op(L, z["u"], float(z["Mass"]), cR, cT, **an)
    assert np.array_equal(op.clov, z["tri"])
    assert _close(op.invclov[:Vh], z["invtri_cb0"], 4e-16)
    assert abs(op.tr_log(0) - z["trlog_cb0"].sum()) < 1e-12 * abs(z["trlog_cb0"].sum())
    for cb in (0, 1):
        assert np.array_equal(oracle.clover_apply(g, z["psi"], z["tri"], cb)[cb * Vh:(cb + 1) * Vh], z["clover_apply_cb%d" % cb])
    chi = np.zeros((2 * Vh, 4, 3, 2)); chi[Vh:] = z["chi"]
    zero = np.zeros_like(chi)
    p, n, res = op.invcg2(chi, zero, 1e-8, 1000)
    assert n == int(z["cg_n"]) and _close(p[Vh:], z["cg_psi"], 1e-12)
    for isign, key in ((+1, "p"), (-1, "m")):
        p, n, res = op.invbicgstab(chi, zero, 1e-8, 1000, isign)
        assert n == int(z["bicg_%s_n" % key]) and _close(p[Vh:], z["bicg_%s_psi" % key], 1e-11)
    p, n = op.minvcg2(chi, z["ms_shifts"], 1e-8, 1000)
    assert n == int(z["ms_n"]) and _close(p[:, Vh:], z["ms_psi"], 1e-11)
    p, n, nupd, _ = op.solve_reliable_cg(chi, zero, 1e-10, 0.1, 1000, mdagm=True)
    assert n == int(z["relcg_n"]) + 1 and _close(p[Vh:], z["relcg_psi"], 1e-11)      # reference counts the zero-based loop index
    p, n, nupd, _ = op.solve_reliable_bicgstab(chi, zero, 1e-10, 0.1, 1000)
    assert n == int(z["relbicg_n"]) + 1 and _close(p[Vh:], z["relbicg_psi"], 1e-8)


@pytest.mark.gpu
@pytest.mark.parametrize("path", CHROMA_FILES, ids=os.path.basename)
def test_cuda_clover_build_matches_chroma_golden(path):
    """a7/a8/a9 on the GPU: b200_make_clover against the reference's mesField + makeClovSiteLoop + LDagDLInvSiteLoop output,
    b200_clover_apply against its applySiteLoop output, log det against its tr_log_diag."""
    from chroma_b200.solver import AnisoParam, CloverFermActParams, Context
    z = np.load(path)
    L, an, cR, cT = _chroma_case(z)
    cp = CloverFermActParams(Mass=float(z["Mass"]), clovCoeffR=cR, clovCoeffT=cT, anisoParam=AnisoParam(**an))
    ctx = Context(L, prec="double")
    ctx.load_gauge(z["u"], aniso_coeff=cp.ferm_coeffs(), t_boundary=-1)
    ctx.make_clover(*cp.derived(), aniso=bool(an), t_dir=3)
    Vh = ctx.Vh
    clov, inv = ctx.get_clover()
    assert np.abs(clov - z["tri"]).max() < 1e-13 * np.abs(z["tri"]).max()
    assert np.abs(inv[:Vh] - z["invtri_cb0"]).max() < 1e-13 * np.abs(z["invtri_cb0"]).max()
    assert abs(ctx.clover_logdet() - z["trlog_cb0"].sum()) < 1e-11 * abs(z["trlog_cb0"].sum())
    for cb in (0, 1):
        got = ctx.clover_apply(z["psi"][cb * Vh:(cb + 1) * Vh], cb)
        assert rel_site_err(got, z["clover_apply_cb%d" % cb]) < 1e-13
    got = ctx.clover_apply(z["psi"][:Vh], 0, inverse=True)
    assert rel_site_err(got, z["invclover_apply_cb0"]) < 1e-12
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("path", CHROMA_FILES, ids=os.path.basename)
def test_cuda_solvers_match_chroma_golden(path):
    """a12/a13/f3/f4 on the GPU against the reference's own compiled solver loops: iteration counts equal (the fused
    reductions sum in another order, so +-1 is allowed) and solutions equal to the accuracy the stopping rule leaves."""
    from chroma_b200 import lib as Lb
    from chroma_b200.solver import AnisoParam, CloverFermActParams, Context
    z = np.load(path)
    L, an, cR, cT = _chroma_case(z)
    cp = CloverFermActParams(Mass=float(z["Mass"]), clovCoeffR=cR, clovCoeffT=cT, anisoParam=AnisoParam(**an))
    ctx = Context(L, prec="double")
    ctx.load_gauge(z["u"], aniso_coeff=cp.ferm_coeffs(), t_boundary=-1)
    ctx.make_clover(*cp.derived(), aniso=bool(an), t_dir=3)
    chi = z["chi"]
    # InvCG2(M, chi, psi) solves M^dag M psi = chi: the MdagM shell (syssolver_mdagm_cg.h:59-94)
    sol, info = ctx.invert_mdagm(chi, None, solver=Lb.B200_SOLVER_CG, rsd=1e-8, max_iter=1000)
    assert abs(info.n_count - int(z["cg_n"])) <= 1 and _close(sol, z["cg_psi"], 1e-6)
    sol, info = ctx.invert(chi, None, solver=Lb.B200_SOLVER_BICGSTAB, rsd=1e-8, max_iter=1000)
    assert abs(info.n_count - int(z["bicg_p_n"])) <= 1 and _close(sol, z["bicg_p_psi"], 1e-6)
    sols, infos = ctx.invert_multishift(chi, z["ms_shifts"], 1e-8, max_iter=1000)
    assert abs(infos[0].n_count - int(z["ms_n"])) <= 1
    for s in range(len(z["ms_shifts"])):
        assert _close(sols[s], z["ms_psi"][s], 1e-6)
    sol, info = ctx.invert_reliable(chi, None, rsd=1e-10, delta=0.1, max_iter=1000, mdagm=True)
    assert abs(info.n_count - (int(z["relcg_n"]) + 1)) <= 2 and _close(sol, z["relcg_psi"], 1e-7)
    sol, info = ctx.invert_reliable_bicgstab(chi, None, rsd=1e-10, delta=0.1, max_iter=1000)
    assert abs(info.n_count - (int(z["relbicg_n"]) + 1)) <= 3 and _close(sol, z["relbicg_psi"], 1e-7)
    ctx.close()
