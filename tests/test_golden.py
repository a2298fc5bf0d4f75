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
