"""Generates tests/golden/ref_*.npz: outputs of the REFERENCE'S OWN code on seeded inputs.

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


if __name__ == "__main__":
    main()
