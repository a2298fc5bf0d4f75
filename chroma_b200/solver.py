"""Host-side mirror of the Chroma plugin for the B200 engine.

`SysSolverB200CloverParams` mirrors the XML group the adapter reads (modelled on SysSolverQUDACloverParams,
lib/actions/ferm/invert/quda_solvers/syssolver_quda_clover_params.cc:15-131) and `LinOpSysSolverB200Clover`
mirrors the plugin class (modelled on LinOpSysSolverQUDAClover, syssolver_linop_clover_quda_w.h:71-648):
construction uploads gauge + clover, `__call__(psi, chi)` solves M psi = chi on the odd checkerboard and returns
SystemSolverResults_t-like (n_count, resid).  Everything numerical happens behind the C ABI (chroma_b200/lib.py).
"""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import lib as L


def _prec_of(a):
    if a.dtype == np.float64:
        return L.B200_DOUBLE
    if a.dtype == np.float32:
        return L.B200_SINGLE
    raise TypeError("host arrays must be float32 or float64, got %s" % a.dtype)


def _ptr(a):
    return C.c_void_p(a.ctypes.data)


class Field:
    """A device-resident checkerboard fermion (b200_field), or a batch of nrhs of them (b200_mfield_alloc)."""

    def __init__(self, ctx, nrhs=1):
        self.ctx = ctx
        self.nrhs = int(nrhs)
        self.h = C.c_void_p()
        if self.nrhs == 1:
            L.check(ctx.lib.b200_field_alloc(ctx.h, C.byref(self.h)))
        else:
            L.check(ctx.lib.b200_mfield_alloc(ctx.h, self.nrhs, C.byref(self.h)))

    def upload(self, host_cb, irhs=None):
        """host_cb: one checkerboard [Vh,4,3,2] (into right-hand side irhs, default 0) or a stack [nrhs,Vh,4,3,2]."""
        host_cb = np.ascontiguousarray(host_cb)
        if irhs is None and host_cb.size == self.nrhs * self.ctx.Vh * 24 and self.nrhs > 1:
            for i in range(self.nrhs):
                self.upload(host_cb.reshape(self.nrhs, -1)[i], i)
            return self
        assert host_cb.size == self.ctx.Vh * 24, "expected one checkerboard [Vh,4,3,2]"
        L.check(self.ctx.lib.b200_mfield_upload(self.ctx.h, self.h, int(irhs or 0), _ptr(host_cb), _prec_of(host_cb)))
        return self

    def download(self, dtype=np.float64, irhs=None):
        """One checkerboard [Vh,4,3,2]; for a batch without irhs, the stack [nrhs,Vh,4,3,2]."""
        if irhs is None and self.nrhs > 1:
            return np.stack([self.download(dtype, i) for i in range(self.nrhs)])
        out = np.empty((self.ctx.Vh, 4, 3, 2), dtype=dtype)
        L.check(self.ctx.lib.b200_mfield_download(self.ctx.h, self.h, int(irhs or 0), _ptr(out), _prec_of(out)))
        return out

    def zero(self):
        L.check(self.ctx.lib.b200_field_zero(self.ctx.h, self.h))
        return self

    def free(self):
        if self.h:
            self.ctx.lib.b200_field_free(self.ctx.h, self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """Thin object wrapper over b200_ctx."""

    def __init__(self, global_dims, prec="double", device=0, proc_grid=(1, 1, 1, 1), proc_coord=(0, 0, 0, 0), comm=None):
        self.lib = L.load()
        self.h = C.c_void_p()
        self.prec = {"double": L.B200_DOUBLE, "single": L.B200_SINGLE, 8: 8, 4: 4}[prec]
        self.global_dims = tuple(int(x) for x in global_dims)
        i4 = C.c_int * 4
        self._comm = comm  # keep callbacks alive
        L.check(self.lib.b200_create(C.byref(self.h), int(device), i4(*self.global_dims), i4(*proc_grid), i4(*proc_coord),
                                     C.byref(comm) if comm is not None else None, self.prec))
        ld = i4()
        L.check(self.lib.b200_local_volume(self.h, ld))
        self.local_dims = tuple(ld)
        self.V = int(np.prod(self.local_dims))
        self.Vh = self.V // 2

    def close(self):
        if self.h:
            self.lib.b200_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- setup
    def load_gauge(self, u, aniso_coeff=(1.0, 1.0, 1.0, 1.0), t_boundary=1, reconstruct=L.B200_RECONS_NONE):
        """u: [4, V, 3, 3, 2] links in QDP++ order with the fermion BC phases already applied."""
        u = np.ascontiguousarray(u)
        assert u.shape[0] == 4 and u[0].size == self.V * 18
        ptrs = (C.c_void_p * 4)(*[u[m].ctypes.data for m in range(4)])
        L.check(self.lib.b200_load_gauge(self.h, ptrs, _prec_of(u), (C.c_double * 4)(*aniso_coeff), int(t_boundary), int(reconstruct)))

    def load_clover(self, clov, invclov):
        clov = np.ascontiguousarray(clov)
        invclov = np.ascontiguousarray(invclov, dtype=clov.dtype)
        assert clov.size == self.V * 72 and invclov.size >= self.Vh * 72
        L.check(self.lib.b200_load_clover(self.h, _ptr(clov), _ptr(invclov), _prec_of(clov)))

    def make_clover(self, diag_mass, clov_r, clov_t, aniso=False, t_dir=3):
        L.check(self.lib.b200_make_clover(self.h, float(diag_mass), float(clov_r), float(clov_t), int(bool(aniso)), int(t_dir)))

    def get_clover(self, dtype=np.float64):
        clov = np.empty((self.V, 72), dtype=dtype)
        inv = np.empty((self.Vh, 72), dtype=dtype)
        L.check(self.lib.b200_get_clover(self.h, _ptr(clov), _ptr(inv), _prec_of(clov)))
        return clov, inv

    def clover_logdet(self, cb=0):
        out = C.c_double()
        fn = self.lib.b200_clover_logdet_oo if cb else self.lib.b200_clover_logdet
        L.check(fn(self.h, C.byref(out)))
        return out.value

    def set_preconditioning(self, symmetric):
        """False: EvenOddPrecCloverLinOp (default); True: SymEvenOddPrecCloverLinOp (seoprec_clover_linop_w.cc:147-193)."""
        L.check(self.lib.b200_set_preconditioning(self.h, L.B200_PRECOND_SYMMETRIC if symmetric else L.B200_PRECOND_ASYMMETRIC))

    def set_twisted_mass(self, mu):
        """CloverFermActParams::twisted_m: M psi += mu i gamma_5 psi (PLUS) / -= (MINUS); 0 = off
        (eoprec_clover_linop_w.cc:174-184, seoprec_clover_linop_w.cc:174-184)."""
        L.check(self.lib.b200_set_twisted_mass(self.h, float(mu)))

    # -- host-buffer operators (what the adapter calls)
    def _cb_out(self, like):
        return np.empty((self.Vh, 4, 3, 2), dtype=like.dtype)

    def dslash(self, in_cb, isign, out_cb):
        in_cb = np.ascontiguousarray(in_cb)
        out = self._cb_out(in_cb)
        L.check(self.lib.b200_dslash(self.h, _ptr(out), _ptr(in_cb), _prec_of(in_cb), int(isign), int(out_cb)))
        return out

    def clover_apply(self, in_cb, cb, inverse=False):
        in_cb = np.ascontiguousarray(in_cb)
        out = self._cb_out(in_cb)
        L.check(self.lib.b200_clover_apply(self.h, _ptr(out), _ptr(in_cb), _prec_of(in_cb), int(cb), int(bool(inverse))))
        return out

    def matpc(self, in_odd, isign=+1):
        in_odd = np.ascontiguousarray(in_odd)
        out = self._cb_out(in_odd)
        L.check(self.lib.b200_clover_matpc(self.h, _ptr(out), _ptr(in_odd), _prec_of(in_odd), int(isign)))
        return out

    def invert(self, chi_odd, psi0_odd=None, solver=L.B200_SOLVER_CG, rsd=1e-8, max_iter=1000):
        chi_odd = np.ascontiguousarray(chi_odd)
        psi = np.zeros_like(chi_odd) if psi0_odd is None else np.ascontiguousarray(psi0_odd, dtype=chi_odd.dtype).copy()
        info = L.SolveInfo()
        rc = self.lib.b200_invert(self.h, _ptr(psi), _ptr(chi_odd), _prec_of(chi_odd), int(solver), float(rsd), int(max_iter), C.byref(info))
        L.check(rc)
        return psi.reshape(self.Vh, 4, 3, 2), info

    def invert_mdagm(self, chi_odd, psi0_odd=None, solver=L.B200_SOLVER_CG, rsd=1e-8, max_iter=1000):
        """Solve M^dag M psi = chi (the HMC-side shells)."""
        chi_odd = np.ascontiguousarray(chi_odd)
        psi = np.zeros_like(chi_odd) if psi0_odd is None else np.ascontiguousarray(psi0_odd, dtype=chi_odd.dtype).copy()
        info = L.SolveInfo()
        L.check(self.lib.b200_invert_mdagm(self.h, _ptr(psi), _ptr(chi_odd), _prec_of(chi_odd), int(solver), float(rsd), int(max_iter), C.byref(info)))
        return psi.reshape(self.Vh, 4, 3, 2), info

    def invert_reliable(self, chi_odd, psi0_odd=None, rsd=1e-8, delta=0.1, max_iter=1000, mdagm=False):
        """Mixed-precision reliable-update CG (fp32 inner, fp64 outer); the context must be double precision."""
        chi_odd = np.ascontiguousarray(chi_odd)
        psi = np.zeros_like(chi_odd) if psi0_odd is None else np.ascontiguousarray(psi0_odd, dtype=chi_odd.dtype).copy()
        info = L.SolveInfo()
        L.check(self.lib.b200_invert_reliable(self.h, _ptr(psi), _ptr(chi_odd), _prec_of(chi_odd), float(rsd), float(delta), int(max_iter),
                                              int(bool(mdagm)), C.byref(info)))
        return psi.reshape(self.Vh, 4, 3, 2), info

    def invert_reliable_bicgstab(self, chi_odd, psi0_odd=None, rsd=1e-8, delta=0.1, max_iter=1000, mdagm=False):
        """Mixed-precision reliable-update BiCGStab (RelInvBiCGStab_a); the context must be double precision."""
        chi_odd = np.ascontiguousarray(chi_odd)
        psi = np.zeros_like(chi_odd) if psi0_odd is None else np.ascontiguousarray(psi0_odd, dtype=chi_odd.dtype).copy()
        info = L.SolveInfo()
        L.check(self.lib.b200_invert_reliable_bicgstab(self.h, _ptr(psi), _ptr(chi_odd), _prec_of(chi_odd), float(rsd), float(delta),
                                                       int(max_iter), int(bool(mdagm)), C.byref(info)))
        return psi.reshape(self.Vh, 4, 3, 2), info

    def invert_multishift(self, chi_odd, shifts, rsd, max_iter=1000):
        """(M^dag M + shifts[s]) psi[s] = chi (MInvCG2_a behind MdagMMultiSysSolverCG).  rsd: scalar or one per shift.
        Returns (psi [n_shift,Vh,4,3,2], [SolveInfo per shift])."""
        chi_odd = np.ascontiguousarray(chi_odd)
        shifts = np.ascontiguousarray(shifts, dtype=np.float64)
        n = len(shifts)
        rsd = np.ascontiguousarray(np.broadcast_to(np.asarray(rsd, dtype=np.float64), (n,)))
        psi = np.zeros((n, self.Vh, 4, 3, 2), dtype=chi_odd.dtype)
        ptrs = (C.c_void_p * n)(*[psi[s].ctypes.data for s in range(n)])
        infos = (L.SolveInfo * n)()
        L.check(self.lib.b200_invert_multishift(self.h, ptrs, _ptr(chi_odd), _prec_of(chi_odd), n,
                                                shifts.ctypes.data_as(C.POINTER(C.c_double)), rsd.ctypes.data_as(C.POINTER(C.c_double)),
                                                int(max_iter), infos))
        return psi, list(infos)

    def qprop(self, chi_full, psi0_full=None, solver=L.B200_SOLVER_CG, rsd=1e-8, max_iter=1000):
        """chi_full: [nrhs, V, 4, 3, 2] full-lattice sources -> full-lattice solutions of the UNPRECONDITIONED operator."""
        chi_full = np.ascontiguousarray(chi_full)
        nrhs = chi_full.shape[0]
        psi = np.zeros_like(chi_full) if psi0_full is None else np.ascontiguousarray(psi0_full, dtype=chi_full.dtype).copy()
        infos = (L.SolveInfo * nrhs)()
        L.check(self.lib.b200_qprop(self.h, _ptr(psi), _ptr(chi_full), _prec_of(chi_full), nrhs, int(solver), float(rsd), int(max_iter), infos))
        return psi, list(infos)

    # -- device-resident operators
    def field(self, host_cb=None):
        f = Field(self)
        if host_cb is not None:
            f.upload(host_cb)
        return f

    def mfield(self, nrhs, host_stack=None):
        """A batch of nrhs checkerboard fermions; host_stack: [nrhs,Vh,4,3,2]."""
        f = Field(self, nrhs)
        if host_stack is not None:
            f.upload(host_stack)
        return f

    def dev_dslash(self, out, inp, isign, out_cb):
        L.check(self.lib.b200_dev_dslash(self.h, out.h, inp.h, int(isign), int(out_cb)))

    def dev_clover_apply(self, out, inp, cb, inverse=False):
        L.check(self.lib.b200_dev_clover_apply(self.h, out.h, inp.h, int(cb), int(bool(inverse))))

    def dev_matpc(self, out, inp, isign=+1):
        L.check(self.lib.b200_dev_clover_matpc(self.h, out.h, inp.h, int(isign)))

    def dev_time_matpc(self, out, inp, isign=+1, reps=10):
        ms = (C.c_double * 2)()
        L.check(self.lib.b200_dev_time_matpc(self.h, out.h, inp.h, int(isign), int(reps), ms))
        return ms[0], ms[1]

    def dev_norm2(self, x):
        r = (C.c_double * x.nrhs)()
        L.check(self.lib.b200_dev_norm2(self.h, x.h, r))
        return r[0] if x.nrhs == 1 else list(r)

    def dev_inner(self, x, y):
        r = (C.c_double * (2 * x.nrhs))()
        L.check(self.lib.b200_dev_inner(self.h, x.h, y.h, r))
        out = [complex(r[2 * i], r[2 * i + 1]) for i in range(x.nrhs)]
        return out[0] if x.nrhs == 1 else out

    def dev_invert(self, psi, chi, solver=L.B200_SOLVER_CG, rsd=1e-8, max_iter=1000):
        """Batched fields: all right-hand sides in lockstep; returns a list of SolveInfo."""
        infos = (L.SolveInfo * psi.nrhs)()
        rc = self.lib.b200_dev_invert(self.h, psi.h, chi.h, int(solver), float(rsd), int(max_iter), infos)
        L.check(rc)
        return infos[0] if psi.nrhs == 1 else list(infos)

    def dev_invert_mdagm(self, psi, chi, solver=L.B200_SOLVER_CG, rsd=1e-8, max_iter=1000):
        infos = (L.SolveInfo * psi.nrhs)()
        L.check(self.lib.b200_dev_invert_mdagm(self.h, psi.h, chi.h, int(solver), float(rsd), int(max_iter), infos))
        return infos[0] if psi.nrhs == 1 else list(infos)

    def dev_invert_reliable(self, psi, chi, rsd=1e-8, delta=0.1, max_iter=1000, mdagm=False):
        info = L.SolveInfo()
        L.check(self.lib.b200_dev_invert_reliable(self.h, psi.h, chi.h, float(rsd), float(delta), int(max_iter), int(bool(mdagm)), C.byref(info)))
        return info

    def dev_invert_reliable_bicgstab(self, psi, chi, rsd=1e-8, delta=0.1, max_iter=1000, mdagm=False):
        info = L.SolveInfo()
        L.check(self.lib.b200_dev_invert_reliable_bicgstab(self.h, psi.h, chi.h, float(rsd), float(delta), int(max_iter),
                                                           int(bool(mdagm)), C.byref(info)))
        return info

    def dev_invert_multishift(self, psi, chi, shifts, rsd, max_iter=1000):
        """psi: a batched field with >= len(shifts) vectors; chi: an ordinary field."""
        shifts = np.ascontiguousarray(shifts, dtype=np.float64)
        n = len(shifts)
        rsd = np.ascontiguousarray(np.broadcast_to(np.asarray(rsd, dtype=np.float64), (n,)))
        infos = (L.SolveInfo * n)()
        L.check(self.lib.b200_dev_invert_multishift(self.h, psi.h, chi.h, n, shifts.ctypes.data_as(C.POINTER(C.c_double)),
                                                    rsd.ctypes.data_as(C.POINTER(C.c_double)), int(max_iter), infos))
        return list(infos)

    def dev_iterate_begin(self, psi, chi, solver):
        L.check(self.lib.b200_dev_iterate_begin(self.h, psi.h, chi.h, int(solver)))

    def dev_iterate(self, solver, n):
        L.check(self.lib.b200_dev_iterate(self.h, int(solver), int(n)))

    def dev_time_solver_kernels(self, solver, reps=5):
        """Average device time (ms) of every launch of one solver iteration, in launch order (after dev_iterate_begin)."""
        ms, n = (C.c_double * 16)(), C.c_int(0)
        L.check(self.lib.b200_dev_time_solver_kernels(self.h, int(solver), int(reps), ms, 16, C.byref(n)))
        return [ms[i] for i in range(n.value)]

    def sync(self):
        L.check(self.lib.b200_sync(self.h))

    @property
    def stream(self):
        return self.lib.b200_stream(self.h)

    @property
    def launch_count(self):
        return int(self.lib.b200_launch_count(self.h))


# ------------------------------------------------------------------------------------------------ plugin mirror
@dataclass
class AnisoParam:
    """AnisoParam_t (lib/io/aniso_io.h)."""
    anisoP: bool = False
    t_dir: int = 3
    xi_0: float = 1.0
    nu: float = 1.0


@dataclass
class CloverFermActParams:
    """CloverFermActParams (lib/actions/ferm/fermacts/clover_fermact_params_w.cc:27-99)."""
    Mass: float = 0.0
    clovCoeffR: float = 1.0
    clovCoeffT: float = 1.0
    anisoParam: AnisoParam = field(default_factory=AnisoParam)
    twisted_m: float = 0.0            # <TwistedM>; twisted_m_usedP = the tag is present (clover_fermact_params_w.cc:90-97)
    twisted_m_usedP: bool = False

    @staticmethod
    def from_kappa(Kappa, clovCoeff, **kw):
        # kappaToMass, lib/io/param_io.cc:12-15
        return CloverFermActParams(Mass=1.0 / (2.0 * Kappa) - 4.0, clovCoeffR=clovCoeff, clovCoeffT=clovCoeff, **kw)

    def derived(self):
        """(diag_mass, clov_r, clov_t) as QDPCloverTermT::create derives them (clover_term_qdp_w.h:263-278)."""
        a = self.anisoParam
        ff = 1.0 / a.xi_0 if a.anisoP else 1.0
        fm = a.nu / a.xi_0 if a.anisoP else 1.0
        return 1.0 + 3.0 * fm + self.Mass, self.clovCoeffR * 0.5 * ff, self.clovCoeffT * 0.5

    def ferm_coeffs(self):
        """makeFermCoeffs (lib/io/aniso_io.cc:63-80)."""
        a = self.anisoParam
        return tuple((a.nu / a.xi_0) if (a.anisoP and mu != a.t_dir) else 1.0 for mu in range(4))


@dataclass
class SysSolverB200CloverParams:
    """The <InvertParam> group of `B200_CLOVER_INVERTER` (schema mirrors syssolver_quda_clover_params.cc:15-131)."""
    CloverParams: CloverFermActParams = field(default_factory=CloverFermActParams)
    RsdTarget: float = 1e-8
    MaxIter: int = 5000
    SolverType: str = "CG"              # CG | BICGSTAB | RELIABLE_CG | RELIABLE_BICGSTAB
    AntiPeriodicT: bool = True
    Precision: str = "DOUBLE"           # SINGLE | DOUBLE  (device precision; XML tag CudaPrecision)
    SloppyPrecision: str = "DEFAULT"    # DEFAULT | SINGLE | DOUBLE; SINGLE under DOUBLE = mixed-precision reliable updates (CudaSloppyPrecision)
    Delta: float = 0.1                  # reliable-update threshold (syssolver_rel_cg_clover_params.cc, syssolver_rel_bicgstab_clover_params.cc:15-18)
    Reconstruct: str = "RECONS_NONE"    # RECONS_NONE | RECONS_12
    RsdToleranceFactor: float = 10.0
    SilentFail: bool = False
    Verbose: bool = False
    # False: the plugin was handed an EvenOddPrecCloverLinOp (clover_fermact_w.cc); True: a SymEvenOddPrecCloverLinOp
    # (seoprec_clover_fermact_w.cc) -- the operator solved must be the caller's A so that its residual check passes
    # (cf. AsymmetricLinop of the QUDA plugin, syssolver_linop_clover_quda_w.h:318-325)
    SymmetricLinop: bool = False
    # True (default): the constructor compares the engine's M with the caller's A on one test vector (checkOperator,
    # chroma_adapter/b200_clover_engine.h); the vector is deterministic, QDP++'s RNG is never touched
    CheckOperator: bool = True


@dataclass
class SystemSolverResults:
    """SystemSolverResults_t (lib/syssolver.h:16-23)."""
    n_count: int = 0
    resid: float = 0.0


class SolverFailure(RuntimeError):
    """Stands in for QDP_abort(1) in the adapter (syssolver_linop_clover_quda_w.h:639-644)."""


class LinOpSysSolverB200Clover:
    """Plugin mirror.  `links` are state->getLinks(): [4,V,3,3,2], fermion BC phases already applied.
    If `clov`/`invclov` (PrimitiveClovTriang arrays) are given they are uploaded (the loadCloverQuda path,
    syssolver_linop_clover_quda_w.h:552); otherwise the GPU builds the clover term from the links."""

    name = "B200_CLOVER_INVERTER"

    def __init__(self, global_dims, links, params: SysSolverB200CloverParams, clov=None, invclov=None, device=0,
                 proc_grid=(1, 1, 1, 1), proc_coord=(0, 0, 0, 0), comm=None, ctx=None):
        self.p = params
        if params.SolverType not in ("CG", "BICGSTAB", "RELIABLE_CG", "RELIABLE_BICGSTAB"):
            raise ValueError("SolverType must be CG, BICGSTAB, RELIABLE_CG or RELIABLE_BICGSTAB")   # adapter: QDPIO::cerr + QDP_abort(1)
        if params.Precision not in ("SINGLE", "DOUBLE"):
            raise ValueError("Precision must be SINGLE or DOUBLE")
        if params.SloppyPrecision not in ("DEFAULT", "SINGLE", "DOUBLE"):
            raise ValueError("SloppyPrecision must be DEFAULT, SINGLE or DOUBLE")
        if params.Reconstruct not in ("RECONS_NONE", "RECONS_12"):
            raise ValueError("Reconstruct must be RECONS_NONE or RECONS_12")
        # as B200CloverEngine's constructor (chroma_adapter/b200_clover_engine.h): mixed precision needs an fp64 engine
        reliable = params.SolverType in ("RELIABLE_CG", "RELIABLE_BICGSTAB")
        if reliable and params.Precision != "DOUBLE":
            raise ValueError("RELIABLE_CG / RELIABLE_BICGSTAB need Precision DOUBLE")
        self.mixed = params.Precision == "DOUBLE" and (params.SloppyPrecision == "SINGLE" or reliable)
        self.bicg = params.SolverType in ("BICGSTAB", "RELIABLE_BICGSTAB")
        self.solver = L.B200_SOLVER_BICGSTAB if self.bicg else L.B200_SOLVER_CG
        self.last_info = None
        if ctx is not None:          # an existing engine context (tests of the dispatch logic; several plugins on one engine)
            self.ctx = ctx
            return
        self.ctx = Context(global_dims, prec=params.Precision.lower(), device=device, proc_grid=proc_grid,
                           proc_coord=proc_coord, comm=comm)
        cp = params.CloverParams
        recon = L.B200_RECONS_12 if params.Reconstruct == "RECONS_12" else L.B200_RECONS_NONE
        self.ctx.load_gauge(links, aniso_coeff=cp.ferm_coeffs(), t_boundary=-1 if params.AntiPeriodicT else 1, reconstruct=recon)
        if clov is not None:
            self.ctx.load_clover(clov, invclov)
        else:
            dm, cr, ct = cp.derived()
            self.ctx.make_clover(dm, cr, ct, aniso=cp.anisoParam.anisoP, t_dir=cp.anisoParam.t_dir)
        if params.SymmetricLinop:
            self.ctx.set_preconditioning(True)
        if cp.twisted_m_usedP:
            self.ctx.set_twisted_mass(cp.twisted_m)

    def subset(self):
        return 1   # rb[1]

    def _solve(self, psi_odd, chi_odd, mdagm):
        """B200CloverEngine::solve (chroma_adapter/b200_clover_engine.h): which ABI entry serves which parameter set."""
        p = self.p
        if self.mixed and not self.bicg:
            return self.ctx.invert_reliable(chi_odd, psi_odd, rsd=p.RsdTarget, delta=p.Delta, max_iter=p.MaxIter, mdagm=mdagm)
        if self.mixed:
            return self.ctx.invert_reliable_bicgstab(chi_odd, psi_odd, rsd=p.RsdTarget, delta=p.Delta, max_iter=p.MaxIter, mdagm=mdagm)
        if mdagm:
            return self.ctx.invert_mdagm(chi_odd, psi_odd, solver=self.solver, rsd=p.RsdTarget, max_iter=p.MaxIter)
        return self.ctx.invert(chi_odd, psi_odd, solver=self.solver, rsd=p.RsdTarget, max_iter=p.MaxIter)

    def __call__(self, psi_odd, chi_odd):
        """psi_odd: initial guess (modified in place), chi_odd: source; both [Vh,4,3,2] on rb[1]."""
        sol, info = self._solve(psi_odd, chi_odd, False)
        psi_odd[...] = sol.reshape(psi_odd.shape)
        self.last_info = info
        res = SystemSolverResults(n_count=info.n_count, resid=info.resid)
        if self.p.Verbose:
            print("B200_%s_CLOVER_SOLVER: %d iterations. Rsd = %g Relative Rsd = %g  time=%g s  Performance=%g GFLOPS"
                  % (self.p.SolverType, info.n_count, info.resid, info.rel_resid, info.secs, info.gflops))
        # the adapter re-verifies with Chroma's own linop and aborts unless SilentFail (.h:634-644)
        if info.rel_resid > self.p.RsdToleranceFactor * self.p.RsdTarget and not self.p.SilentFail:
            raise SolverFailure("B200 solver failed to converge: rel resid %g > %g * %g" %
                                (info.rel_resid, self.p.RsdToleranceFactor, self.p.RsdTarget))
        return res

    def close(self):
        self.ctx.close()


class MdagMMultiSysSolverB200Clover(LinOpSysSolverB200Clover):
    """Mirror of the multi-shift plugin: MdagMMultiSysSolverCG::operator()(psi[], shifts, chi)
    (lib/actions/ferm/invert/multi_syssolver_mdagm_cg.h:58-105), registered in TheMdagMFermMultiSystemSolverFactory
    (multi_syssolver_mdagm_factory.h) under the same `B200_CLOVER_INVERTER` name.  RsdTarget applies to every shift
    (multi_syssolver_mdagm_cg.h:62-70 broadcasts a single RsdCG)."""

    def __call__(self, shifts, chi_odd):
        psi, infos = self.ctx.invert_multishift(chi_odd, shifts, self.p.RsdTarget, self.p.MaxIter)
        self.last_info = infos
        if not infos[0].converged and not self.p.SilentFail:
            raise SolverFailure("B200 multi-shift CG: too many iterations (%d)" % infos[0].n_count)   # minvcg2.cc:365-367
        worst = max(i.rel_resid for i in infos)
        if worst > self.p.RsdToleranceFactor * self.p.RsdTarget and not self.p.SilentFail:
            raise SolverFailure("B200 multi-shift CG: rel resid %g > %g * %g" % (worst, self.p.RsdToleranceFactor, self.p.RsdTarget))
        return psi, SystemSolverResults(n_count=infos[0].n_count, resid=max(i.resid for i in infos))


class MdagMSysSolverB200Clover(LinOpSysSolverB200Clover):
    """Mirror of the HMC-side plugin (chroma_adapter/syssolver_mdagm_clover_b200_w.h; twin of MdagMSysSolverQUDAClover):
    M^dag M psi = chi, with the optional chronological-predictor overload of MdagMSysSolverCG
    (syssolver_mdagm_cg.h:104-123): `predictor(psi, chi)` supplies the initial guess, `predictor.new_vector(psi)`
    records the solution (AbsChronologicalPredictor4D::operator() / newVector)."""

    def __call__(self, psi_odd, chi_odd, predictor=None):
        if predictor is not None:
            predictor(psi_odd, chi_odd)
        sol, info = self._solve(psi_odd, chi_odd, True)
        psi_odd[...] = sol.reshape(psi_odd.shape)
        self.last_info = info
        if self.p.Verbose:
            print("B200_CLOVER_SOLVER (MdagM): %d iterations. Rsd = %g Relative Rsd = %g" % (info.n_count, info.resid, info.rel_resid))
        if info.rel_resid > self.p.RsdToleranceFactor * self.p.RsdTarget and not self.p.SilentFail:
            raise SolverFailure("B200 MdagM solver residuum is outside tolerance: rel resid %g > %g * %g" %
                                (info.rel_resid, self.p.RsdToleranceFactor, self.p.RsdTarget))
        if predictor is not None:
            predictor.new_vector(psi_odd)
        return SystemSolverResults(n_count=info.n_count, resid=info.resid)
