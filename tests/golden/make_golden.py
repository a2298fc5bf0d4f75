"""Generates tests/golden/ref_*.npz and tests/golden/chroma_*.npz: outputs of the REFERENCE'S OWN code on seeded inputs.

Run in the build container only (needs /root/reference, through oracle/_ref/libref_dslash.so which oracle/Makefile
compiles, unmodified and in place, from /root/reference/other_libs/cpp_wilson_dslash/lib):

    OMP_NUM_THREADS=1 python tests/golden/make_golden.py

What is recorded (inputs AND outputs, so the fixtures do not depend on any RNG staying bit-stable):
  * Dslash<double>::operator() and Dslash<float>::operator() (cpp_dslash_scalar_64bit.cc:35-65, ..._32bit.cc) for
    isign = +-1 and source checkerboard 0/1, on random SU(3) links with an antiperiodic T boundary -- the procedure of
    the reference's tests/testDslashFull.cc:70-97 (which holds no stored vectors: SURVEY.md section 8c);
  * CloverSchur4D<double>::operator() (cpp_clover_scalar_64bit.cc:65-380) for isign = +-1 with links x 1/2, a clover
    term built by the restated makeClov/ldagdlinv (Mass 0.1, c_sw 1.0) and Im(offd[0][14]) zeroed (the one entry where
    the reference's plain-C cloverSiteApply is not Hermitian, tests/test_oracle.py::test_reference_clover_site_apply_defect).

The fixtures pin the oracle restatement (tests/test_golden.py, CPU) and the CUDA kernels (tests/test_golden.py -m gpu)
to the reference without /root/reference being present.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from chroma_b200 import fields  # noqa: E402
from oracle import oracle as orc  # noqa: E402

CASES = [("4x4x4x4", (4, 4, 4, 4), 101, True), ("6x4x2x4", (6, 4, 2, 4), 201, False)]


def main():
    assert os.environ.get("OMP_NUM_THREADS") == "1", "run with OMP_NUM_THREADS=1 (CloverSchur4D has no barrier between its site loops)"
    orc.build()
    assert orc.have_ref(), "oracle/_ref/libref_dslash.so missing: run `make -C oracle` where /root/reference exists"
    for name, L, seed, with_clover in CASES:
        g = orc.Geom(L)
        u = fields.apply_bc(L, fields.random_gauge(L, seed=seed))
        psi = fields.gaussian_fermion(L, seed=seed + 1)
        pk = orc.pack_gauge(L, u, (1.0, 1.0, 1.0, 1.0))
        out = {"L": np.array(L, dtype=np.int32), "u": u, "psi": psi}
        rd, rf = orc.RefDslash(L, np.float64), orc.RefDslash(L, np.float32)
        for isign in (+1, -1):
            for cb in (0, 1):
                tgt = slice((1 - cb) * g.Vh, (2 - cb) * g.Vh)
                key = "%s_cb%d" % ("p" if isign > 0 else "m", cb)
                out["dslash_d_" + key] = rd(psi, pk, isign, cb)[tgt]
                out["dslash_f_" + key] = rf(psi.astype(np.float32), pk.astype(np.float32), isign, cb)[tgt]
        if with_clover:
            op = orc.Op(L, u, 0.1, 1.0)
            clov, invclov = op.clov.copy(), op.invclov.copy()
            clov[:, 41] = 0.0
            invclov[:, 41] = 0.0
            op.clov[:, 41] = 0.0
            op.invclov[:, 41] = 0.0
            half = orc.pack_gauge(L, u, (0.5, 0.5, 0.5, 0.5))
            ref = orc.RefCloverSchur(L)
            chi = psi.copy()
            chi[:g.Vh] = 0.0
            out["clov"] = clov
            out["invclov_ee"] = invclov[:g.Vh]
            for isign in (+1, -1):
                out["schur_d_" + ("p" if isign > 0 else "m")] = ref(chi, half, orc.tri_to_ref_clover(clov),
                                                                    orc.tri_to_ref_clover(invclov), isign)[g.Vh:]
        path = os.path.join(HERE, "ref_%s.npz" % name)
        np.savez(path, **out)
        print(path, "%.0f KB" % (os.path.getsize(path) / 1024))


def main_chroma():
    """tests/golden/chroma_*.npz: outputs of the reference's Chroma-level code for the path -- mesField, the clover site
    loops of clover_term_qdp_w.h and the solver loops invcg2 / invbicgstab / minvcg2 / reliable_cg / reliable_bicgstab --
    compiled unmodified into oracle/_ref/libref_chroma.so (oracle/Makefile).  The operator the solvers iterate on is the
    oracle's (pinned to the reference Dslash / CloverSchur4D by the ref_*.npz fixtures above)."""
    assert orc.have_ref_chroma(), "oracle/_ref/libref_chroma.so missing: run `make -C oracle` where /root/reference exists"
    for name, L, seed, aniso in (("4x4x4x8", (4, 4, 4, 8), 301, False), ("6x4x4x4_aniso", (6, 4, 4, 4), 401, True)):
        g = orc.Geom(L)
        u = fields.apply_bc(L, fields.weak_gauge(L, seed=seed))
        an = dict(anisoP=True, t_dir=3, xi_0=2.464, nu=0.95) if aniso else {}
        cR, cT = (0.91, 1.07) if aniso else (1.0, 1.0)
        dm, r, t = orc.clover_coeffs(0.1, cR, cT, **{k: v for k, v in an.items() if k != "t_dir"})
        f = orc.ref_mesfield(L, u)
        tri = orc.ref_make_clov(L, f, dm, r, t, anisoP=aniso, t_dir=3)
        inv0, trlog0 = orc.ref_ldagdlinv(L, tri, 0)
        inv1, trlog1 = orc.ref_ldagdlinv(L, tri, 1)
        psi = fields.gaussian_fermion(L, seed=seed + 1)
        out = {"L": np.array(L, dtype=np.int32), "u": u, "Mass": 0.1, "clovCoeffR": cR, "clovCoeffT": cT, "aniso": int(aniso),
               "xi_0": an.get("xi_0", 1.0), "nu": an.get("nu", 1.0),
               "f": f, "tri": tri, "invtri_cb0": inv0[:g.Vh], "invtri_cb1": inv1[g.Vh:], "trlog_cb0": trlog0[:g.Vh], "trlog_cb1": trlog1[g.Vh:],
               "psi": psi, "clover_apply_cb0": orc.ref_clover_apply(L, psi, tri, 0)[:g.Vh],
               "clover_apply_cb1": orc.ref_clover_apply(L, psi, tri, 1)[g.Vh:],
               "invclover_apply_cb0": orc.ref_clover_apply(L, psi, inv0, 0)[:g.Vh]}
        op = orc.Op(L, u, 0.1, cR, cT, **an)
        chi = fields.gaussian_fermion(L, seed=seed + 2, cb=1)
        zero = np.zeros_like(chi)
        out["chi"] = chi[g.Vh:]
        p, n, res, tr = orc.ref_invcg2(op, chi, zero, 1e-8, 1000)
        out.update(cg_psi=p[g.Vh:], cg_n=n, cg_resid=res, cg_trace=tr)
        for isign, key in ((+1, "p"), (-1, "m")):
            p, n, res, tr = orc.ref_invbicgstab(op, chi, zero, 1e-8, 1000, isign)
            out.update({"bicg_%s_psi" % key: p[g.Vh:], "bicg_%s_n" % key: n, "bicg_%s_resid" % key: res, "bicg_%s_trace" % key: tr})
        shifts = np.array([0.7, 0.001, 0.05])      # unsorted on purpose: minvcg2.cc:153-164 finds the smallest itself
        p, n, tr = orc.ref_minvcg2(op, chi, shifts, 1e-8, 1000)
        out.update(ms_shifts=shifts, ms_psi=p[:, g.Vh:], ms_n=n, ms_trace=tr)
        p, n, res, c64, c32 = orc.ref_reliable_cg(op, chi, zero, 1e-10, 0.1, 1000)
        out.update(relcg_psi=p[g.Vh:], relcg_n=n, relcg_resid=res, relcg_calls=np.array([c64, c32]))
        p, n, res, c64, c32 = orc.ref_reliable_bicgstab(op, chi, zero, 1e-10, 0.1, 1000)
        out.update(relbicg_psi=p[g.Vh:], relbicg_n=n, relbicg_resid=res, relbicg_calls=np.array([c64, c32]))
        path = os.path.join(HERE, "chroma_%s.npz" % name)
        np.savez_compressed(path, **out)
        print(path, "%.0f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
    main_chroma()
