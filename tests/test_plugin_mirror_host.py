"""CPU tests of the host-side plugin mirror (chroma_b200/solver.py): which C-ABI entry each XML parameter set is routed
to (the logic of B200CloverEngine::solve in chroma_adapter/b200_clover_engine.h), parameter validation, the failure
policy (QDP_abort unless SilentFail, syssolver_linop_clover_quda_w.h:634-644) and the chronological-predictor overload
(syssolver_mdagm_cg.h:104-123).  The engine context is a recording stand-in: nothing here computes."""
import numpy as np
import pytest

from chroma_b200 import lib as L
from chroma_b200.solver import (CloverFermActParams, AnisoParam, LinOpSysSolverB200Clover, MdagMMultiSysSolverB200Clover,
                                MdagMSysSolverB200Clover, SolverFailure, SysSolverB200CloverParams)


class Info:
    def __init__(self, n_count=7, resid=1e-9, rel_resid=1e-9, converged=1):
        self.n_count, self.resid, self.rel_resid, self.converged = n_count, resid, rel_resid, converged
        self.secs = self.gflops = 0.0


class RecordingContext:
    def __init__(self, info=None):
        self.calls = []
        self.info = info or Info()

    def _ret(self, name, chi, psi, **kw):
        self.calls.append((name, kw))
        return np.full_like(chi, 2.0), self.info

    def invert(self, chi, psi, **kw): return self._ret("invert", chi, psi, **kw)
    def invert_mdagm(self, chi, psi, **kw): return self._ret("invert_mdagm", chi, psi, **kw)
    def invert_reliable(self, chi, psi, **kw): return self._ret("invert_reliable", chi, psi, **kw)
    def invert_reliable_bicgstab(self, chi, psi, **kw): return self._ret("invert_reliable_bicgstab", chi, psi, **kw)

    def invert_multishift(self, chi, shifts, rsd, max_iter):
        self.calls.append(("invert_multishift", dict(shifts=list(shifts), rsd=rsd, max_iter=max_iter)))
        return np.zeros((len(shifts),) + chi.shape), [self.info for _ in shifts]


def make(cls, ctx, **kw):
    return cls((4, 4, 4, 4), None, SysSolverB200CloverParams(**kw), ctx=ctx)


@pytest.mark.parametrize("kw,mdagm,entry,solver", [
    (dict(SolverType="CG"), False, "invert", L.B200_SOLVER_CG),
    (dict(SolverType="BICGSTAB"), False, "invert", L.B200_SOLVER_BICGSTAB),
    (dict(SolverType="CG"), True, "invert_mdagm", L.B200_SOLVER_CG),
    (dict(SolverType="BICGSTAB"), True, "invert_mdagm", L.B200_SOLVER_BICGSTAB),
    (dict(SolverType="RELIABLE_CG"), False, "invert_reliable", None),
    (dict(SolverType="CG", SloppyPrecision="SINGLE"), True, "invert_reliable", None),
    (dict(SolverType="RELIABLE_BICGSTAB"), True, "invert_reliable_bicgstab", None),
    (dict(SolverType="BICGSTAB", SloppyPrecision="SINGLE"), False, "invert_reliable_bicgstab", None),
    (dict(SolverType="BICGSTAB", SloppyPrecision="SINGLE", Precision="SINGLE"), False, "invert", L.B200_SOLVER_BICGSTAB),
])
def test_parameter_sets_reach_the_right_abi_entry(kw, mdagm, entry, solver):
    ctx = RecordingContext()
    S = make(MdagMSysSolverB200Clover if mdagm else LinOpSysSolverB200Clover, ctx, RsdTarget=1e-7, MaxIter=123, Delta=0.05, **kw)
    psi, chi = np.zeros((8, 4, 3, 2)), np.ones((8, 4, 3, 2))
    res = S(psi, chi)
    assert res.n_count == 7 and res.resid == 1e-9 and np.all(psi == 2.0)           # solution written into the caller's psi
    (name, args), = ctx.calls
    assert name == entry
    if solver is None:
        assert args == dict(rsd=1e-7, delta=0.05, max_iter=123, mdagm=mdagm)
    else:
        assert args == dict(solver=solver, rsd=1e-7, max_iter=123)


def test_parameter_validation_and_failure_policy():
    for bad in (dict(SolverType="GCR"), dict(Precision="HALF"), dict(SloppyPrecision="HALF"), dict(Reconstruct="RECONS_8"),
                dict(SolverType="RELIABLE_CG", Precision="SINGLE")):
        with pytest.raises(ValueError):
            make(LinOpSysSolverB200Clover, RecordingContext(), **bad)
    psi, chi = np.zeros((8, 4, 3, 2)), np.ones((8, 4, 3, 2))
    far = RecordingContext(Info(rel_resid=1e-3))
    with pytest.raises(SolverFailure):
        make(LinOpSysSolverB200Clover, far, RsdTarget=1e-8)(psi, chi)
    assert make(LinOpSysSolverB200Clover, far, RsdTarget=1e-8, SilentFail=True)(psi, chi).n_count == 7
    assert make(LinOpSysSolverB200Clover, far, RsdTarget=1e-8, RsdToleranceFactor=1e6)(psi, chi).n_count == 7
    with pytest.raises(SolverFailure):
        make(MdagMSysSolverB200Clover, far, RsdTarget=1e-8)(psi, chi)


def test_chronological_predictor_overload():
    class Predictor:
        def __init__(self): self.log = []
        def __call__(self, psi, chi): self.log.append("guess"); psi[...] = 0.5
        def new_vector(self, psi): self.log.append(("new", float(psi.flat[0])))
    pred = Predictor()
    S = make(MdagMSysSolverB200Clover, RecordingContext())
    psi, chi = np.zeros((8, 4, 3, 2)), np.ones((8, 4, 3, 2))
    S(psi, chi, pred)
    assert pred.log == ["guess", ("new", 2.0)]


def test_multishift_mirror_policy():
    psi_chi = np.ones((8, 4, 3, 2))
    ctx = RecordingContext()
    S = make(MdagMMultiSysSolverB200Clover, ctx, RsdTarget=1e-6, MaxIter=50)
    psi, res = S([0.1, 0.2, 0.3], psi_chi)
    assert psi.shape == (3, 8, 4, 3, 2) and res.n_count == 7
    assert ctx.calls == [("invert_multishift", dict(shifts=[0.1, 0.2, 0.3], rsd=1e-6, max_iter=50))]
    with pytest.raises(SolverFailure):
        make(MdagMMultiSysSolverB200Clover, RecordingContext(Info(converged=0)), RsdTarget=1e-6)([0.1], psi_chi)
    with pytest.raises(SolverFailure):
        make(MdagMMultiSysSolverB200Clover, RecordingContext(Info(rel_resid=1.0)), RsdTarget=1e-6)([0.1], psi_chi)


def test_clover_parameter_derivations():
    """QDPCloverTermT::create (clover_term_qdp_w.h:263-278), makeFermCoeffs (io/aniso_io.cc:63-80), kappaToMass
    (io/param_io.cc:12-15)."""
    p = CloverFermActParams(Mass=0.1, clovCoeffR=0.91, clovCoeffT=1.07, anisoParam=AnisoParam(anisoP=True, t_dir=3, xi_0=2.464, nu=0.95))
    dm, cr, ct = p.derived()
    assert dm == pytest.approx(1.0 + 3.0 * 0.95 / 2.464 + 0.1) and cr == pytest.approx(0.5 * 0.91 / 2.464) and ct == pytest.approx(0.535)
    assert p.ferm_coeffs() == pytest.approx((0.95 / 2.464,) * 3 + (1.0,))
    iso = CloverFermActParams.from_kappa(0.115, 1.27)
    assert iso.Mass == pytest.approx(1.0 / 0.23 - 4.0) and iso.derived() == pytest.approx((1.0 + 3.0 + iso.Mass, 0.635, 0.635))
    assert iso.ferm_coeffs() == (1.0, 1.0, 1.0, 1.0)


def test_constructor_sequence_with_a_stand_in_engine(monkeypatch):
    """The constructor's call sequence on the engine (as LinOpSysSolverQUDAClover's ctor: create, load gauge with the
    anisotropy factors and the boundary sign, build or load the clover term, select the preconditioning), with
    chroma_b200.solver.Context replaced by a recorder."""
    from chroma_b200 import solver as S

    class FakeContext(RecordingContext):
        def __init__(self, global_dims, prec="double", device=0, proc_grid=(1, 1, 1, 1), proc_coord=(0, 0, 0, 0), comm=None):
            super().__init__()
            self.calls.append(("create", dict(dims=tuple(global_dims), prec=prec, device=device, grid=tuple(proc_grid))))

        def load_gauge(self, u, aniso_coeff, t_boundary, reconstruct):
            self.calls.append(("load_gauge", dict(aniso=tuple(aniso_coeff), t_boundary=t_boundary, reconstruct=reconstruct)))

        def make_clover(self, dm, cr, ct, aniso=False, t_dir=3):
            self.calls.append(("make_clover", dict(dm=dm, cr=cr, ct=ct, aniso=aniso, t_dir=t_dir)))

        def load_clover(self, clov, invclov):
            self.calls.append(("load_clover", {}))

        def set_preconditioning(self, sym):
            self.calls.append(("set_preconditioning", dict(sym=sym)))

        def close(self):
            self.calls.append(("close", {}))

    monkeypatch.setattr(S, "Context", FakeContext)
    cp = CloverFermActParams(Mass=0.1, clovCoeffR=1.0, clovCoeffT=1.0)
    p = SysSolverB200CloverParams(CloverParams=cp, RsdTarget=1e-8, MaxIter=1000, SolverType="BICGSTAB", AntiPeriodicT=True,
                                  Reconstruct="RECONS_12", SymmetricLinop=True)
    sol = LinOpSysSolverB200Clover((8, 8, 8, 8), "links", p, device=3)
    names = [c[0] for c in sol.ctx.calls]
    assert names == ["create", "load_gauge", "make_clover", "set_preconditioning"]
    assert sol.ctx.calls[0][1] == dict(dims=(8, 8, 8, 8), prec="double", device=3, grid=(1, 1, 1, 1))
    assert sol.ctx.calls[1][1] == dict(aniso=(1.0, 1.0, 1.0, 1.0), t_boundary=-1, reconstruct=L.B200_RECONS_12)
    assert sol.ctx.calls[2][1] == dict(dm=4.1, cr=0.5, ct=0.5, aniso=False, t_dir=3)
    psi, chi = np.zeros((8, 4, 3, 2)), np.ones((8, 4, 3, 2))
    assert sol(psi, chi).n_count == 7 and sol.ctx.calls[-1] == ("invert", dict(solver=L.B200_SOLVER_BICGSTAB, rsd=1e-8, max_iter=1000))
    sol.close()
    assert sol.ctx.calls[-1][0] == "close"
    # handing over Chroma's own clover buffers instead (the loadCloverQuda path); periodic T; fp32 engine
    p2 = SysSolverB200CloverParams(CloverParams=cp, AntiPeriodicT=False, Precision="SINGLE")
    m = MdagMMultiSysSolverB200Clover((4, 4, 4, 4), "links", p2, clov="clov", invclov="invclov")
    assert [c[0] for c in m.ctx.calls] == ["create", "load_gauge", "load_clover"]
    assert m.ctx.calls[0][1]["prec"] == "single" and m.ctx.calls[1][1]["t_boundary"] == 1
    h = MdagMSysSolverB200Clover((4, 4, 4, 4), "links", SysSolverB200CloverParams(CloverParams=cp, SolverType="RELIABLE_CG"))
    assert h(psi, chi).n_count == 7 and h.ctx.calls[-1][0] == "invert_reliable" and h.ctx.calls[-1][1]["mdagm"] is True
