"""ctypes/numpy front-end of the CPU oracle (oracle/oracle.c) and of the reference build (oracle/_ref).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / reference arm.  Nothing under chroma_b200/ imports this module.

Array conventions (QDP++ order):
  spinor  float64[V, 4, 3, 2]      gauge  float64[4, V, 3, 3, 2]      clover  float64[V, 72]
  site index = cb2: cb*Vh + ((t*Lz+z)*Ly+y)*(Lx/2) + x/2
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

c_int4 = C.c_int * 4
c_dbl_p = C.POINTER(C.c_double)
c_flt_p = C.POINTER(C.c_float)


def build(force=False):
    """Compile liboracle.so (and oracle/_ref when /root/reference is present)."""
    if force or not os.path.exists(os.path.join(_HERE, "liboracle.so")) or \
            os.path.getmtime(os.path.join(_HERE, "liboracle.so")) < os.path.getmtime(os.path.join(_HERE, "oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/other_libs/cpp_wilson_dslash/lib"):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


def _p(a):
    return a.ctypes.data_as(c_dbl_p)


def lib():
    global _LIB
    if _LIB is None:
        build()
        L = C.CDLL(os.path.join(_HERE, "liboracle.so"))
        L.orc_geom_create.restype = C.c_void_p
        L.orc_op_create.restype = C.c_void_p
        L.orc_op_geom.restype = C.c_void_p
        L.orc_op_clov.restype = c_dbl_p
        L.orc_op_invclov.restype = c_dbl_p
        L.orc_op_packed_gauge.restype = c_dbl_p
        L.orc_norm2_odd.restype = C.c_double
        L.orc_op_tr_log.restype = C.c_double
        _LIB = L
    return _LIB


def have_ref():
    return os.path.exists(os.path.join(_HERE, "_ref", "libref_dslash.so"))


def ref():
    """The reference's own Dslash<double|float> / CloverSchur4D<double> (oracle/_ref/libref_dslash.so)."""
    global _REF
    if _REF is None:
        R = C.CDLL(os.path.join(_HERE, "_ref", "libref_dslash.so"))
        for n in ("ref_dslash_create_d", "ref_dslash_create_f", "ref_clover_create_d"):
            getattr(R, n).restype = C.c_void_p
        _REF = R
    return _REF


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    """OpenMP team size of the CPU arm (torchrun exports OMP_NUM_THREADS=1; bench.py asks for every core it may use)."""
    lib().orc_set_num_threads(C.c_int(int(n)))


def site_index(L, c):
    return lib().orc_site_index(c_int4(*L), c_int4(*c))


def site_coords_all(L):
    """int array [V,4] of (x,y,z,t) for every cb2 site index (vectorised restatement of orc_site_coords)."""
    Lx, Ly, Lz, Lt = L
    V = Lx * Ly * Lz * Lt
    Vh, Lxh = V // 2, Lx // 2
    idx = np.arange(V)
    cb, r = idx // Vh, idx % Vh
    xh = r % Lxh
    r //= Lxh
    y = r % Ly
    r //= Ly
    z = r % Lz
    t = r // Lz
    x = 2 * xh + ((cb + y + z + t) & 1)
    return np.stack([x, y, z, t], axis=1)


class Geom:
    def __init__(self, L):
        self.L = tuple(L)
        self.V = int(np.prod(L))
        self.Vh = self.V // 2
        self.h = C.c_void_p(lib().orc_geom_create(c_int4(*L)))

    def __del__(self):
        try:
            lib().orc_geom_free(self.h)
        except Exception:
            pass


def apply_fermbc(L, u, boundary):
    u = np.ascontiguousarray(u, dtype=np.float64).copy()
    ptrs = (c_dbl_p * 4)(*[_p(u[m]) for m in range(4)])
    lib().orc_apply_fermbc(c_int4(*L), ptrs, c_int4(*boundary))
    return u


def pack_gauge(L, u, coeffs=(1.0, 1.0, 1.0, 1.0)):
    V = int(np.prod(L))
    u = np.ascontiguousarray(u, dtype=np.float64)
    out = np.empty((V, 4, 3, 3, 2), dtype=np.float64)
    ptrs = (c_dbl_p * 4)(*[_p(u[m]) for m in range(4)])
    lib().orc_pack_gauge(c_int4(*L), ptrs, (C.c_double * 4)(*coeffs), _p(out))
    return out


def dslash(geom, psi, packed_u, isign, source_cb):
    """Restated Dslash<double>::operator(): writes the (1-source_cb) half, other half zero."""
    psi = np.ascontiguousarray(psi, dtype=np.float64)
    res = np.zeros_like(psi)
    lib().orc_dslash(geom.h, _p(res), _p(psi), _p(packed_u), C.c_int(isign), C.c_int(source_cb))
    return res


def mesfield(geom, u):
    u = np.ascontiguousarray(u, dtype=np.float64)
    f = np.empty((6, geom.V, 3, 3, 2), dtype=np.float64)
    ptrs = (c_dbl_p * 4)(*[_p(u[m]) for m in range(4)])
    lib().orc_mesfield(geom.h, ptrs, _p(f))
    return f


def clover_coeffs(Mass, clovCoeffR, clovCoeffT, anisoP=False, xi_0=1.0, nu=1.0):
    out = (C.c_double * 3)()
    lib().orc_clover_coeffs(C.c_double(Mass), C.c_double(clovCoeffR), C.c_double(clovCoeffT), C.c_int(int(anisoP)),
                            C.c_double(xi_0), C.c_double(nu), out)
    return tuple(out)


def make_clov(geom, f, diag_mass, cR, cT, anisoP=False, t_dir=3):
    tri = np.empty((geom.V, 72), dtype=np.float64)
    lib().orc_make_clov(geom.h, _p(f), C.c_double(diag_mass), C.c_double(cR), C.c_double(cT), C.c_int(int(anisoP)),
                        C.c_int(t_dir), _p(tri))
    return tri


def ldagdlinv(geom, tri, cb):
    out = np.ascontiguousarray(tri, dtype=np.float64).copy()
    trlog = np.zeros(geom.V, dtype=np.float64)
    lib().orc_ldagdlinv(geom.h, _p(out), C.c_int(cb), _p(trlog))
    return out, trlog


def clover_apply(geom, psi, tri, cb):
    psi = np.ascontiguousarray(psi, dtype=np.float64)
    chi = np.zeros_like(psi)
    lib().orc_clover_apply(geom.h, _p(chi), _p(psi), _p(tri), C.c_int(cb))
    return chi


def norm2_odd(geom, x):
    return lib().orc_norm2_odd(geom.h, _p(np.ascontiguousarray(x)))


class Op:
    """Restated EvenOddPrecCloverLinOp (eoprec_clover_linop_w.cc:19-39, 142-187).
    u: float64[4,V,3,3,2] links WITH fermion boundary phases already applied."""

    def __init__(self, L, u, Mass, clovCoeffR, clovCoeffT=None, anisoP=False, t_dir=3, xi_0=1.0, nu=1.0):
        if clovCoeffT is None:
            clovCoeffT = clovCoeffR
        self.L = tuple(L)
        self.V = int(np.prod(L))
        self.Vh = self.V // 2
        u = np.ascontiguousarray(u, dtype=np.float64)
        ptrs = (c_dbl_p * 4)(*[_p(u[m]) for m in range(4)])
        self.h = C.c_void_p(lib().orc_op_create(c_int4(*L), ptrs, C.c_double(Mass), C.c_double(clovCoeffR),
                                                C.c_double(clovCoeffT), C.c_int(int(anisoP)), C.c_int(t_dir),
                                                C.c_double(xi_0), C.c_double(nu)))
        self.geom_h = C.c_void_p(lib().orc_op_geom(self.h))

    def __del__(self):
        try:
            lib().orc_op_free(self.h)
        except Exception:
            pass

    def _arr(self, ptr, shape):
        return np.ctypeslib.as_array(ptr, shape=shape)

    def set_symmetric(self, sym=True):
        """Switch to SymEvenOddPrecCloverLinOp (seoprec_clover_linop_w.cc:16-41, 147-193): 1 - 1/4 A_oo^-1 D A_ee^-1 D.
        apply(), the solvers and the qprop prepare / reconstruct steps all follow the switch."""
        lib().orc_op_set_symmetric(self.h, C.c_int(int(bool(sym))))

    def set_twisted_mass(self, mu):
        """CloverFermActParams::twisted_m: apply() then ends with chi +/-= mu * Gamma(15) * timesI(psi)
        (eoprec_clover_linop_w.cc:174-184, seoprec_clover_linop_w.cc:174-184); 0 switches the term off."""
        lib().orc_op_set_twisted_mass(self.h, C.c_double(float(mu)))

    @property
    def symmetric(self):
        return bool(lib().orc_op_is_symmetric(self.h))

    def tr_log(self, cb):
        """sum over checkerboard cb of log|det A| (cb 1 only after set_symmetric)."""
        return float(lib().orc_op_tr_log(self.h, C.c_int(cb)))

    def minvcg2(self, chi, shifts, rsd, maxit):
        """MInvCG2_a (minvcg2.cc:74-373): (M^dag M + shifts[s]) psi[s] = chi.  Returns (psi[n_shift], n_count)."""
        shifts = np.ascontiguousarray(shifts, dtype=np.float64)
        rsd = np.ascontiguousarray(np.broadcast_to(np.asarray(rsd, dtype=np.float64), shifts.shape))
        chi = np.ascontiguousarray(chi, dtype=np.float64)
        psi = np.zeros((len(shifts),) + chi.shape)
        n = lib().orc_minvcg2(self.h, _p(chi), _p(psi), _p(shifts), _p(rsd), C.c_int(len(shifts)), C.c_int(maxit))
        return psi, n

    def solve_multishift(self, chi, shifts, rsd, maxit):
        """MdagMMultiSysSolverCG::operator() (multi_syssolver_mdagm_cg.h:58-105): MInvCG2 and the per-shift relative
        residuals |chi - (M^dag M + shift) psi| / |chi| it logs."""
        psi, n = self.minvcg2(chi, shifts, rsd, maxit)
        Vh = self.Vh
        cn = np.sqrt(np.sum(np.asarray(chi)[Vh:] ** 2))
        rel = []
        for s, sh in enumerate(shifts):
            r = chi - self.apply(self.apply(psi[s], +1), -1) - sh * psi[s]
            rel.append(float(np.sqrt(np.sum(r[Vh:] ** 2)) / cn))
        return psi, n, rel

    @property
    def clov(self):
        return self._arr(lib().orc_op_clov(self.h), (self.V, 72))

    @property
    def invclov(self):
        return self._arr(lib().orc_op_invclov(self.h), (self.V, 72))

    @property
    def packed_gauge(self):
        return self._arr(lib().orc_op_packed_gauge(self.h), (self.V, 4, 3, 3, 2))

    def dslash(self, psi, isign, target_cb):
        psi = np.ascontiguousarray(psi, dtype=np.float64)
        chi = np.zeros_like(psi)
        lib().orc_op_dslash(self.h, _p(chi), _p(psi), C.c_int(isign), C.c_int(target_cb))
        return chi

    def apply(self, psi, isign=+1):
        psi = np.ascontiguousarray(psi, dtype=np.float64)
        chi = np.zeros_like(psi)
        lib().orc_op_apply(self.h, _p(chi), _p(psi), C.c_int(isign))
        return chi

    def unprec_apply(self, psi, isign=+1):
        psi = np.ascontiguousarray(psi, dtype=np.float64)
        chi = np.zeros_like(psi)
        lib().orc_unprec_apply(self.h, _p(chi), _p(psi), C.c_int(isign))
        return chi

    def qprop_prepare(self, chi):
        chi = np.ascontiguousarray(chi, dtype=np.float64)
        out = np.zeros_like(chi)
        lib().orc_qprop_prepare(self.h, _p(out), _p(chi))
        return out

    def qprop_reconstruct(self, psi, chi):
        psi = np.ascontiguousarray(psi, dtype=np.float64).copy()
        lib().orc_qprop_reconstruct(self.h, _p(psi), _p(np.ascontiguousarray(chi)))
        return psi

    def invcg2(self, chi, psi0, rsd, maxit):
        psi = np.ascontiguousarray(psi0, dtype=np.float64).copy()
        resid = C.c_double()
        n = lib().orc_invcg2(self.h, _p(np.ascontiguousarray(chi)), _p(psi), C.c_double(rsd), C.c_int(maxit), C.byref(resid))
        return psi, n, resid.value

    def solve_cg(self, chi, psi0, rsd, maxit):
        psi = np.ascontiguousarray(psi0, dtype=np.float64).copy()
        out = (C.c_double * 2)()
        n = lib().orc_solve_cg(self.h, _p(np.ascontiguousarray(chi)), _p(psi), C.c_double(rsd), C.c_int(maxit), out)
        return psi, n, out[0], out[1]

    def solve_bicgstab(self, chi, psi0, rsd, maxit):
        psi = np.ascontiguousarray(psi0, dtype=np.float64).copy()
        out = (C.c_double * 2)()
        n = lib().orc_solve_bicgstab(self.h, _p(np.ascontiguousarray(chi)), _p(psi), C.c_double(rsd), C.c_int(maxit), out)
        return psi, n, out[0], out[1]

    def invbicgstab(self, chi, psi0, rsd, maxit, isign=+1):
        """InvBiCGStab_a with an explicit isign (invbicgstab.cc:10-202)."""
        psi = np.ascontiguousarray(psi0, dtype=np.float64).copy()
        resid = C.c_double()
        n = lib().orc_invbicgstab(self.h, _p(np.ascontiguousarray(chi)), _p(psi), C.c_double(rsd), C.c_int(maxit), C.c_int(isign), C.byref(resid))
        return psi, n, resid.value

    def _mdagm_resid(self, chi, psi):
        r = chi - self.apply(self.apply(psi, +1), -1)
        return float(np.sqrt(np.sum(r[self.Vh:] ** 2)))

    def solve_mdagm_cg(self, chi, psi0, rsd, maxit):
        """MdagMSysSolverCG::operator() (syssolver_mdagm_cg.h:59-94): InvCG2 on chi itself; resid = |chi - M^dag M psi|."""
        psi, n, _ = self.invcg2(chi, psi0, rsd, maxit)
        return psi, n, self._mdagm_resid(chi, psi)

    def solve_mdagm_bicgstab(self, chi, psi0, rsd, maxit):
        """MdagMSysSolverBiCGStab::operator() (syssolver_mdagm_bicgstab.h:62-110): Y = M psi; M^dag Y = chi; M psi = Y."""
        Y = self.apply(psi0, +1)
        Y, n1, _ = self.invbicgstab(chi, Y, rsd, maxit, -1)
        psi, n2, _ = self.invbicgstab(Y, psi0, rsd, maxit, +1)
        return psi, n1 + n2, self._mdagm_resid(chi, psi)

    def solve_reliable_cg(self, chi, psi0, rsd, delta, maxit, mdagm=False):
        """RelInvCG_a<double, float> (lib/actions/ferm/invert/reliable_cg.cc:10-190) behind the CGNE shell of
        LinOpSysSolverReliableCGClover (syssolver_linop_rel_cg_clover.h:118-150).  The fp32 operator AF is emulated
        as round_to_float(M(float vector)) with M evaluated in double -- the rounding of the vectors, which is what
        drives the reliable-update logic, is exact; the internal rounding of an fp32 Dslash is not modelled.
        Returns (psi, iterations [1-based count; the reference reports its 0-based loop index], n_updates, resid)."""
        f32 = lambda a: a.astype(np.float32).astype(np.float64)   # noqa: E731
        Vh = self.Vh
        n2 = lambda a: float(np.sum(a[Vh:] ** 2))                 # noqa: E731
        rhs = chi if mdagm else self.apply(chi, -1)
        psi = np.ascontiguousarray(psi0, dtype=np.float64).copy()
        chi_norm = n2(rhs)
        rsd_sq = rsd * rsd * chi_norm
        b = np.zeros_like(psi)
        b[Vh:] = (rhs - self.apply(self.apply(psi, +1), -1))[Vh:]
        x = np.zeros_like(psi)
        r = f32(b)
        r_sq = n2(r)
        rNorm = np.sqrt(r_sq)
        r0Norm = maxrx = maxrr = rNorm
        p = np.zeros_like(psi)
        c = 1.0
        n_upd = 0
        iters = maxit
        for k in range(maxit):
            if k == 0:
                p = r.copy()
            else:
                beta = np.float32(r_sq / c)
                p = f32(r + float(beta) * p)
            c = r_sq
            mp = f32(self.apply(p, +1))
            d = n2(mp)
            mmp = f32(self.apply(mp, -1))
            a = float(np.float32(c / d))
            x = f32(x + a * p)
            r = f32(r - a * mmp)
            r_sq = n2(r)
            rNorm = np.sqrt(r_sq)
            maxrx = max(maxrx, rNorm)
            maxrr = max(maxrr, rNorm)
            updateX = rNorm < delta * r0Norm and r0Norm <= maxrx
            updateR = (rNorm < delta * maxrr and r0Norm <= maxrr) or updateX
            if updateR:
                n_upd += 1
                r_d = np.zeros_like(psi)
                r_d[Vh:] = (b - self.apply(self.apply(x, +1), -1))[Vh:]
                r = f32(r_d)
                r_sq = n2(r_d)
                rNorm = np.sqrt(r_sq)
                maxrr = rNorm
                if updateX:
                    psi[Vh:] += x[Vh:]
                    x[:] = 0
                    b = r_d
                    r0Norm = maxrx = rNorm
            if r_sq < rsd_sq:
                psi[Vh:] += x[Vh:]
                iters = k + 1
                break
        else:
            psi[Vh:] += x[Vh:]
        resid = self._mdagm_resid(chi, psi) if mdagm else float(np.sqrt(n2(chi - self.apply(psi, +1))))
        return psi, iters, n_upd, resid


    def solve_reliable_bicgstab(self, chi, psi0, rsd, delta, maxit, isign=+1):
        """RelInvBiCGStab_a<double, float> (lib/actions/ferm/invert/reliable_bicgstab.cc:13-290) behind the shell of
        LinOpSysSolverReliableBiCGStabClover (syssolver_linop_rel_bicgstab_clover.h:105-146): A psi = chi with A = M
        (isign=+1) or M^dag (-1).  fp32 vectors emulated by rounding, as in solve_reliable_cg.  Note that, like the
        reference, rho is NOT recomputed after a residual replacement (:186, :205-216).
        Returns (psi, iterations [1-based], n_r_updates, |chi - A psi|)."""
        Vh = self.Vh
        c64 = lambda a: (a[..., 0] + 1j * a[..., 1])                                     # noqa: E731
        f32 = lambda z: z.astype(np.complex64).astype(np.complex128)                     # noqa: E731

        def A(z):          # complex odd-cb vector -> complex odd-cb vector, evaluated in double
            full = np.zeros((self.V, 4, 3, 2))
            full[Vh:, ..., 0] = z.real
            full[Vh:, ..., 1] = z.imag
            return c64(self.apply(full, isign)[Vh:])

        chi_c = c64(np.asarray(chi, dtype=np.float64)[Vh:])
        psi = c64(np.ascontiguousarray(psi0, dtype=np.float64)[Vh:]).copy()
        rsd_sq = rsd * rsd * float(np.vdot(chi_c, chi_c).real)
        b = chi_c - A(psi)
        r_sq = float(np.vdot(b, b).real)
        r = f32(b)
        r0 = r.copy()
        x = np.zeros_like(r)
        p = np.zeros_like(r)
        v = np.zeros_like(r)
        rNorm = np.sqrt(r_sq)
        r0Norm = maxrx = maxrr = rNorm
        rho = rho_prev = alpha = omega = 1.0 + 0j
        n_upd = 0
        iters = maxit
        for k in range(maxit):
            if k == 0:
                rho = r_sq + 0j
                p = r.copy()
            else:
                beta = np.complex64((rho / rho_prev) * (alpha / omega))
                p = f32(r + complex(beta) * f32(p - complex(np.complex64(omega)) * v))
            v = f32(A(p))
            ctmp = np.vdot(r0, v)
            if ctmp == 0:
                raise ArithmeticError("BiCGStab breakdown: <r_0|v> = 0")
            alpha = rho / ctmp
            rho_prev = rho
            r = f32(r - complex(np.complex64(alpha)) * v)
            t = f32(A(r))
            omega = np.vdot(t, r) / float(np.vdot(t, t).real)
            x = f32(f32(x + complex(np.complex64(omega)) * r) + complex(np.complex64(alpha)) * p)
            r = f32(r - complex(np.complex64(omega)) * t)
            r_sq = float(np.vdot(r, r).real)
            rho = np.vdot(r0, r)
            rNorm = np.sqrt(r_sq)
            maxrx = max(maxrx, rNorm)
            maxrr = max(maxrr, rNorm)
            updateX = rNorm < delta * r0Norm and r0Norm <= maxrx
            updateR = (rNorm < delta * maxrr and r0Norm <= maxrr) or updateX
            if updateR:
                n_upd += 1
                r_d = b - A(x)
                r_sq = float(np.vdot(r_d, r_d).real)
                r = f32(r_d)
                rNorm = np.sqrt(r_sq)
                maxrr = rNorm
                if updateX:
                    psi = psi + x
                    x = np.zeros_like(x)
                    b = r_d
                    r0Norm = maxrx = rNorm
            if r_sq < rsd_sq:
                psi = psi + x
                iters = k + 1
                break
        else:
            psi = psi + x
        out = np.ascontiguousarray(psi0, dtype=np.float64).copy()
        out[Vh:, ..., 0] = psi.real
        out[Vh:, ..., 1] = psi.imag
        res = chi_c - A(psi)
        return out, iters, n_upd, float(np.sqrt(np.vdot(res, res).real))

    def solve_mdagm_reliable_bicgstab(self, chi, psi0, rsd, delta, maxit):
        """MdagMSysSolverReliableBiCGStabClover::operator() (syssolver_mdagm_rel_bicgstab_clover.h:104-170): Y = M psi;
        M^dag Y = chi (MINUS); M psi = Y (PLUS); n_count = sum; resid = |chi - M^dag M psi|."""
        Y = self.apply(psi0, +1)
        Y, n1, u1, _ = self.solve_reliable_bicgstab(chi, Y, rsd, delta, maxit, -1)
        psi, n2, u2, _ = self.solve_reliable_bicgstab(Y, psi0, rsd, delta, maxit, +1)
        return psi, n1 + n2, u1 + u2, self._mdagm_resid(chi, psi)


# --------------------------------------------------------------------------- reference build
class RefDslash:
    """Reference CPlusPlusWilsonDslash::Dslash<double|float> (cpp_dslash_scalar.h:20-105)."""

    def __init__(self, L, dtype=np.float64):
        self.dtype = np.dtype(dtype)
        self.sfx = "d" if self.dtype == np.float64 else "f"
        self.h = C.c_void_p(getattr(ref(), "ref_dslash_create_" + self.sfx)(c_int4(*L)))

    def __call__(self, psi, packed_u, isign, source_cb):
        psi = np.ascontiguousarray(psi, dtype=self.dtype)
        u = np.ascontiguousarray(packed_u, dtype=self.dtype)
        res = np.zeros_like(psi)
        getattr(ref(), "ref_dslash_apply_" + self.sfx)(self.h, C.c_void_p(res.ctypes.data), C.c_void_p(psi.ctypes.data),
                                                        C.c_void_p(u.ctypes.data), C.c_int(isign), C.c_int(source_cb))
        return res

    def __del__(self):
        try:
            getattr(ref(), "ref_dslash_free_" + self.sfx)(self.h)
        except Exception:
            pass


def tri_to_ref_clover(tri):
    """Chroma PrimitiveClovTriang [V,72] -> cpp_clover CloverTerm [V,2,{diag[8],off_diag[16][2]}] (80 reals),
    same k = i(i-1)/2+j order (cpp_clover_types.h:8-27)."""
    V = tri.shape[0]
    out = np.zeros((V, 2, 40), dtype=tri.dtype)
    for b in range(2):
        out[:, b, 0:6] = tri[:, 6 * b:6 * b + 6]
        out[:, b, 8:38] = tri[:, 12 + 30 * b:12 + 30 * b + 30]
    return out


class RefCloverSchur:
    """Reference CPlusPlusClover::CloverSchur4D<double> (cpp_clover_scalar.h). Run with OMP_NUM_THREADS=1
    (its two site loops are not separated by a barrier)."""

    def __init__(self, L):
        self.h = C.c_void_p(ref().ref_clover_create_d(c_int4(*L)))

    def __call__(self, psi, packed_u, clov80, invclov80, isign):
        psi = np.ascontiguousarray(psi, dtype=np.float64)
        res = np.zeros_like(psi)
        ref().ref_clover_apply_d(self.h, _p(res), _p(psi), _p(np.ascontiguousarray(packed_u)),
                                 _p(np.ascontiguousarray(clov80)), _p(np.ascontiguousarray(invclov80)), C.c_int(isign))
        return res

    def __del__(self):
        try:
            ref().ref_clover_free_d(self.h)
        except Exception:
            pass


def cg_bench(op, chi, n_warm, n_timed, use_reference_dslash=True):
    """Seconds per CG iteration / per M apply on the host cores (bench.py's CPU arm).  With use_reference_dslash the
    hopping term is the reference's own Dslash<double> from oracle/_ref; clover apply and BLAS are the restatement
    (they live in QDP++-dependent files that cannot be compiled here)."""
    kind = "port"
    keep = None
    if use_reference_dslash and have_ref():
        rd = RefDslash(op.L, np.float64)
        fn = ref().ref_dslash_apply_d
        lib().orc_set_dslash_hook(C.cast(fn, C.c_void_p), rd.h)
        kind, keep = "reference", rd
    else:
        lib().orc_set_dslash_hook(None, None)
    out = (C.c_double * 2)()
    lib().orc_cg_bench(op.h, _p(np.ascontiguousarray(chi)), C.c_int(n_warm), C.c_int(n_timed), out)
    lib().orc_set_dslash_hook(None, None)
    del keep
    return out[0], out[1], kind


# ------------------------------------------------------------------------------------------------------------------
# The reference's Chroma-level code for the path, compiled unmodified into oracle/_ref/libref_chroma.so against
# tests/mock_chroma (oracle/Makefile, oracle/ref_chroma_shim.cc): clover site loops of clover_term_qdp_w.h and the solver
# loops invcg2.cc / invbicgstab.cc / minvcg2.cc / reliable_cg.cc / reliable_bicgstab.cc.  Used by tests/test_oracle.py to
# pin the restatements above and by tests/golden/make_golden.py to write fixtures.
_REFC = None


def have_ref_chroma():
    return os.path.exists(os.path.join(_HERE, "_ref", "libref_chroma.so"))


def ref_chroma(L):
    """dlopen the library and set its (global, single-rank) layout to lattice L."""
    global _REFC
    if _REFC is None:
        _REFC = C.CDLL(os.path.join(_HERE, "_ref", "libref_chroma.so"))
    _REFC.refc_setup(c_int4(*L))
    return _REFC


def _apply_ptr():
    """orc_op_apply(op, chi, psi, isign) as the C callback the reference solvers iterate on."""
    return C.cast(lib().orc_op_apply, C.c_void_p)


def ref_mesfield(L, u):
    """The reference's mesField (lib/meas/glue/mesfield.cc:30-78) on a periodic single-rank lattice."""
    R = ref_chroma(L)
    u = np.ascontiguousarray(u, dtype=np.float64)
    f = np.zeros((6,) + u.shape[1:], dtype=np.float64)
    R.refc_mesfield(_p(u), _p(f))
    return f


def ref_make_clov(L, f, diag_mass, cR, cT, anisoP=False, t_dir=3):
    """QDPCloverTermT::makeClov with the reference's own makeClovSiteLoop (clover_term_qdp_w.h:398-553)."""
    R = ref_chroma(L)
    pmu, pnu = (0, 0, 0, 1, 1, 2), (1, 2, 3, 2, 3, 3)
    coef = (C.c_double * 6)(*[(cT if (anisoP and (pmu[k] == t_dir or pnu[k] == t_dir)) else cR) for k in range(6)])
    f = np.ascontiguousarray(f, dtype=np.float64)
    tri = np.zeros((f.shape[1], 72), dtype=np.float64)
    R.refc_make_clov(_p(f), coef, C.c_double(diag_mass), _p(tri))
    return tri


def ref_ldagdlinv(L, tri, cb):
    """QDPCloverTermT::ldagdlinv with the reference's own LDagDLInvSiteLoop (:619-846)."""
    R = ref_chroma(L)
    out = np.ascontiguousarray(tri, dtype=np.float64).copy()
    trlog = np.zeros(out.shape[0], dtype=np.float64)
    R.refc_ldagdlinv(_p(out), C.c_int(cb), _p(trlog))
    return out, trlog


def ref_clover_apply(L, psi, tri, cb):
    """QDPCloverTermT::apply with the reference's own applySiteLoop (:1562-1634)."""
    R = ref_chroma(L)
    psi = np.ascontiguousarray(psi, dtype=np.float64)
    chi = np.zeros_like(psi)
    R.refc_clover_apply(_p(np.ascontiguousarray(tri, dtype=np.float64)), _p(psi), _p(chi), C.c_int(cb))
    return chi


def ref_invcg2(op, chi, psi0, rsd, maxit, trace_cap=4096):
    """The reference's InvCG2 on the oracle operator `op`.  Returns (psi, n_count, resid, trace) with trace[i] = |input|^2 of
    the i-th operator application -- an iteration-by-iteration fingerprint of the Krylov sequence."""
    R = ref_chroma(op.L)
    psi = np.ascontiguousarray(psi0, dtype=np.float64).copy()
    out, trace = (C.c_double * 4)(), np.zeros(trace_cap)
    R.refc_invcg2(_apply_ptr(), op.h, _p(np.ascontiguousarray(chi, dtype=np.float64)), _p(psi), C.c_double(rsd), C.c_int(maxit),
                  out, _p(trace), C.c_int(trace_cap))
    return psi, int(out[0]), out[1], trace[:min(int(out[2]), trace_cap)]


def ref_invbicgstab(op, chi, psi0, rsd, maxit, isign=+1, trace_cap=4096):
    R = ref_chroma(op.L)
    psi = np.ascontiguousarray(psi0, dtype=np.float64).copy()
    out, trace = (C.c_double * 4)(), np.zeros(trace_cap)
    R.refc_invbicgstab(_apply_ptr(), op.h, _p(np.ascontiguousarray(chi, dtype=np.float64)), _p(psi), C.c_double(rsd),
                       C.c_int(maxit), C.c_int(isign), out, _p(trace), C.c_int(trace_cap))
    return psi, int(out[0]), out[1], trace[:min(int(out[2]), trace_cap)]


def ref_minvcg2(op, chi, shifts, rsd, maxit, trace_cap=4096):
    R = ref_chroma(op.L)
    shifts = np.ascontiguousarray(shifts, dtype=np.float64)
    rsd = np.ascontiguousarray(np.broadcast_to(np.asarray(rsd, dtype=np.float64), shifts.shape))
    chi = np.ascontiguousarray(chi, dtype=np.float64)
    psi = np.zeros((len(shifts),) + chi.shape)
    out, trace = (C.c_double * 4)(), np.zeros(trace_cap)
    R.refc_minvcg2(_apply_ptr(), op.h, _p(chi), _p(psi), _p(shifts), _p(rsd), C.c_int(len(shifts)), C.c_int(maxit), out,
                   _p(trace), C.c_int(trace_cap))
    return psi, int(out[0]), trace[:min(int(out[2]), trace_cap)]


def ref_reliable_cg(op, chi, psi0, rsd, delta, maxit):
    """The reference's InvCGReliable(A, AF, ...) with the oracle operator in fp64 and the same operator rounded through
    float fields as the fp32 one.  Returns (psi, n_count, resid, fp64 applications, fp32 applications)."""
    R = ref_chroma(op.L)
    psi = np.ascontiguousarray(psi0, dtype=np.float64).copy()
    out = (C.c_double * 4)()
    R.refc_reliable_cg(_apply_ptr(), op.h, _p(np.ascontiguousarray(chi, dtype=np.float64)), _p(psi), C.c_double(rsd),
                       C.c_double(delta), C.c_int(maxit), out)
    return psi, int(out[0]), out[1], int(out[2]), int(out[3])


def ref_reliable_bicgstab(op, chi, psi0, rsd, delta, maxit, isign=+1):
    R = ref_chroma(op.L)
    psi = np.ascontiguousarray(psi0, dtype=np.float64).copy()
    out = (C.c_double * 4)()
    R.refc_reliable_bicgstab(_apply_ptr(), op.h, _p(np.ascontiguousarray(chi, dtype=np.float64)), _p(psi), C.c_double(rsd),
                             C.c_double(delta), C.c_int(maxit), C.c_int(isign), out)
    return psi, int(out[0]), out[1], int(out[2]), int(out[3])
