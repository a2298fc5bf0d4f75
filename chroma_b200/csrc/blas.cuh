// blas.cuh -- solver BLAS fused with deterministic grid reductions; scalars stay on the device.
//
// The fusions mirror what Chroma's CPU path hand-fuses in
// lib/actions/ferm/invert/bicgstab_kernels_scalarsite.h:30-344 (xmay_normx_cdotzx, yxpaymabz,
// norm2x_cdotxy, xpaypbz, cxmay, xymz_normx) and the axpy/norm2 sequence of invcg2.cc:174-220, but
// go further: the CG residual update lives in the Dslash epilogue (dslash.cuh, EPI_M_CG) so that
// M^dag M p is never written to HBM.
//
// All kernels run over the flat array of n = 12*Vh complex numbers with a FIXED grid, thread t of
// block b handling elements b*BLOCK+t, +GRID*BLOCK, ... so partial sums are bitwise reproducible.  The hot loops are
// unrolled by 4 so that a thread has the loads of four elements in flight (the two-stream r -= alpha v ran at 0.67 of the
// HBM peak with one).
#pragma once
#include "common.cuh"
#include "reduce.cuh"

namespace b200 {

constexpr int BLAS_BLOCK = 256;

// Batched (multi-RHS) launches use gridDim.y = number of right-hand sides: blockIdx.y picks the field (fstride
// elements apart), the scalar / status block and the reduction slots of that right-hand side.  gridDim.y = 1 is the
// ordinary single-RHS launch.
struct BlasCtl {
  double* scal; int* status; ReduceBuf red;
  int iter; int check_stop;
  size_t fstride;
  int rel;        // bicg_update_kernel: also take the reliable-update decisions (reliable_bicgstab.cc:198-205)
  template <int N>
  __device__ __forceinline__ BlasCtl for_rhs(int rhs) const {
    BlasCtl c = *this;
    c.scal += rhs * S_COUNT; c.status += rhs * ST_COUNT;
    if (N > 0) c.red = red.template for_rhs<(N > 0 ? N : 1)>(rhs);
    return c;
  }
};

// ---------------------------------------------------------------- CG: psi += a p ; p = r + b p
// invcg2.cc:185 and :220.  On the converging iteration psi is still updated, p is not.
template <typename R>
__global__ void __launch_bounds__(BLAS_BLOCK) cg_update_kernel(Cx<R>* __restrict__ psi, Cx<R>* __restrict__ p,
                                                              const Cx<R>* __restrict__ r, size_t n, BlasCtl c0) {
  const BlasCtl c = c0.template for_rhs<0>(blockIdx.y);
  psi += blockIdx.y * c.fstride; p += blockIdx.y * c.fstride; r += blockIdx.y * c.fstride;
  int stop = c.check_stop ? c.status[ST_STOP] : 0;
  if (stop != 0 && stop < c.iter) return;
  const bool conv = (stop == c.iter) && stop != 0;
  const R a = (R)c.scal[S_A], b = (R)c.scal[S_B];
#pragma unroll 4
  for (size_t i = (size_t)blockIdx.x * BLAS_BLOCK + threadIdx.x; i < n; i += (size_t)gridDim.x * BLAS_BLOCK) {
    Cx<R> pv = p[i], xv = psi[i];
    xv.x += a * pv.x; xv.y += a * pv.y;
    psi[i] = xv;
    if (!conv) {
      const Cx<R> rv = r[i];
      pv.x = rv.x + b * pv.x; pv.y = rv.y + b * pv.y;
      p[i] = pv;
    }
  }
}

// ---------------------------------------------------------------- BiCGStab: p = r + beta (p - omega v)
// invbicgstab.cc:87-97
template <typename R>
__global__ void __launch_bounds__(BLAS_BLOCK) bicg_p_kernel(Cx<R>* __restrict__ p, const Cx<R>* __restrict__ r,
                                                           const Cx<R>* __restrict__ v, size_t n, BlasCtl c0) {
  const BlasCtl c = c0.template for_rhs<0>(blockIdx.y);
  p += blockIdx.y * c.fstride; r += blockIdx.y * c.fstride; v += blockIdx.y * c.fstride;
  if (c.check_stop && (c.status[ST_STOP] != 0 || c.status[ST_BREAKDOWN] != 0)) return;
  const Cx<R> beta = mk<R>((R)c.scal[S_BETA_RE], (R)c.scal[S_BETA_IM]);
  const Cx<R> omega = mk<R>((R)c.scal[S_OMEGA_RE], (R)c.scal[S_OMEGA_IM]);
#pragma unroll 4
  for (size_t i = (size_t)blockIdx.x * BLAS_BLOCK + threadIdx.x; i < n; i += (size_t)gridDim.x * BLAS_BLOCK) {
    const Cx<R> tmp = csub(p[i], cmul(omega, v[i]));
    p[i] = cadd(r[i], cmul(beta, tmp));
  }
}

// ---------------------------------------------------------------- BiCGStab: r -= alpha v   (s overlaps r)
// invbicgstab.cc:122
template <typename R>
__global__ void __launch_bounds__(BLAS_BLOCK) bicg_s_kernel(Cx<R>* __restrict__ r, const Cx<R>* __restrict__ v, size_t n, BlasCtl c0) {
  const BlasCtl c = c0.template for_rhs<0>(blockIdx.y);
  r += blockIdx.y * c.fstride; v += blockIdx.y * c.fstride;
  if (c.check_stop && (c.status[ST_STOP] != 0 || c.status[ST_BREAKDOWN] != 0)) return;
  const Cx<R> alpha = mk<R>((R)c.scal[S_ALPHA_RE], (R)c.scal[S_ALPHA_IM]);
#pragma unroll 4
  for (size_t i = (size_t)blockIdx.x * BLAS_BLOCK + threadIdx.x; i < n; i += (size_t)gridDim.x * BLAS_BLOCK)
    r[i] = csub(r[i], cmul(alpha, v[i]));
}

// ---------------------------------------------------------------- BiCGStab: psi += omega r + alpha p ; r -= omega t ;
// |r|^2 and rho_next = <r0|r> in the same pass (invbicgstab.cc:149-160 and :77 of the next iteration)
struct FinBiUpdate {
  double* scal; int* status; int iter; int check; int rel;
  __device__ void operator()(const double* t) const {
    const double rnorm = t[0], nr = t[1], ni = t[2];
    scal[S_RNORM] = rnorm;
    bool conv = false, upd_r = false;
    if (rel) {
      // reliable updates (reliable_bicgstab.cc:193-205): decide here, on the device, whether the fp64 side replaces the
      // residual (updateR) and folds the fp32 partial solution into psi (updateX); if so the convergence test moves to
      // the finaliser of the replacement kernel (FinRelReplace), which sees the TRUE residual
      const double rn = sqrt(rnorm);
      double maxrx = scal[S_MAXRX], maxrr = scal[S_MAXRR];
      const double r0 = scal[S_R0NORM], delta = scal[S_DELTA];
      if (rn > maxrx) maxrx = rn;
      if (rn > maxrr) maxrr = rn;
      scal[S_MAXRX] = maxrx; scal[S_MAXRR] = maxrr;
      const bool upd_x = (rn < delta * r0) && (r0 <= maxrx);
      upd_r = ((rn < delta * maxrr) && (r0 <= maxrr)) || upd_x;
      status[ST_UPD_R] = upd_r ? 1 : 0; status[ST_UPD_X] = upd_x ? 1 : 0;
    }
    if (!upd_r && check && status[ST_STOP] == 0 && rnorm < scal[S_RSDSQ]) { status[ST_STOP] = iter; conv = true; }
    const double pr = scal[S_RHO_RE], pi = scal[S_RHO_IM];
    scal[S_RHOP_RE] = pr; scal[S_RHOP_IM] = pi;
    scal[S_RHO_RE] = nr; scal[S_RHO_IM] = ni;
    if (nr == 0.0 && ni == 0.0) { if (!conv && status[ST_BREAKDOWN] == 0) status[ST_BREAKDOWN] = 1; return; }
    // beta = (rho/rho_prev) * (alpha/omega)   (invbicgstab.cc:87)
    const double d1 = pr * pr + pi * pi;
    const double qr = (nr * pr + ni * pi) / d1, qi = (ni * pr - nr * pi) / d1;
    const double ar = scal[S_ALPHA_RE], ai = scal[S_ALPHA_IM], wr = scal[S_OMEGA_RE], wi = scal[S_OMEGA_IM];
    const double d2 = wr * wr + wi * wi;
    const double sr = (ar * wr + ai * wi) / d2, si = (ai * wr - ar * wi) / d2;
    scal[S_BETA_RE] = qr * sr - qi * si;
    scal[S_BETA_IM] = qr * si + qi * sr;
  }
};

template <typename R>
__global__ void __launch_bounds__(BLAS_BLOCK) bicg_update_kernel(Cx<R>* __restrict__ psi, Cx<R>* __restrict__ r,
                                                                const Cx<R>* __restrict__ p, const Cx<R>* __restrict__ t,
                                                                const Cx<R>* __restrict__ r0, size_t n, BlasCtl c0) {
  const BlasCtl c = c0.template for_rhs<3>(blockIdx.y);
  psi += blockIdx.y * c.fstride; r += blockIdx.y * c.fstride; p += blockIdx.y * c.fstride; t += blockIdx.y * c.fstride; r0 += blockIdx.y * c.fstride;
  if (c.check_stop && (c.status[ST_STOP] != 0 || c.status[ST_BREAKDOWN] != 0)) return;
  const Cx<R> alpha = mk<R>((R)c.scal[S_ALPHA_RE], (R)c.scal[S_ALPHA_IM]);
  const Cx<R> omega = mk<R>((R)c.scal[S_OMEGA_RE], (R)c.scal[S_OMEGA_IM]);
  double red[3] = {0.0, 0.0, 0.0};
#pragma unroll 4
  for (size_t i = (size_t)blockIdx.x * BLAS_BLOCK + threadIdx.x; i < n; i += (size_t)gridDim.x * BLAS_BLOCK) {
    Cx<R> rv = r[i];
    const Cx<R> tmp = cadd(psi[i], cmul(omega, rv));
    psi[i] = cadd(tmp, cmul(alpha, p[i]));
    rv = csub(rv, cmul(omega, t[i]));
    r[i] = rv;
    const Cx<R> q = r0[i];
    red[0] += (double)rv.x * rv.x + (double)rv.y * rv.y;
    red[1] += (double)q.x * rv.x + (double)q.y * rv.y;
    red[2] += (double)q.x * rv.y - (double)q.y * rv.x;
  }
  grid_reduce<3, BLAS_BLOCK>(red, c.red, FinBiUpdate{c.scal, c.status, c.iter, c.check_stop, c.rel});
}

// ---------------------------------------------------------------- generic helpers (setup / verification, not the hot loop)
struct FinStore { double* dst; int n; __device__ void operator()(const double* t) const { for (int k = 0; k < n; ++k) dst[k] = t[k]; } };

template <typename R>
__global__ void __launch_bounds__(BLAS_BLOCK) norm2_kernel(const Cx<R>* __restrict__ x, size_t n, ReduceBuf red0, double* dst, size_t fstride) {
  const ReduceBuf red = red0.template for_rhs<1>(blockIdx.y);
  x += blockIdx.y * fstride; dst += blockIdx.y * S_COUNT;
  double s[1] = {0.0};
  for (size_t i = (size_t)blockIdx.x * BLAS_BLOCK + threadIdx.x; i < n; i += (size_t)gridDim.x * BLAS_BLOCK) {
    const Cx<R> v = x[i]; s[0] += (double)v.x * v.x + (double)v.y * v.y;
  }
  grid_reduce<1, BLAS_BLOCK>(s, red, FinStore{dst, 1});
}

template <typename R>
__global__ void __launch_bounds__(BLAS_BLOCK) inner_kernel(const Cx<R>* __restrict__ x, const Cx<R>* __restrict__ y, size_t n,
                                                          ReduceBuf red0, double* dst, size_t fstride) {
  const ReduceBuf red = red0.template for_rhs<2>(blockIdx.y);
  x += blockIdx.y * fstride; y += blockIdx.y * fstride; dst += blockIdx.y * S_COUNT;
  double s[2] = {0.0, 0.0};
  for (size_t i = (size_t)blockIdx.x * BLAS_BLOCK + threadIdx.x; i < n; i += (size_t)gridDim.x * BLAS_BLOCK) {
    const Cx<R> a = x[i], b = y[i];
    s[0] += (double)a.x * b.x + (double)a.y * b.y;
    s[1] += (double)a.x * b.y - (double)a.y * b.x;
  }
  grid_reduce<2, BLAS_BLOCK>(s, red, FinStore{dst, 2});
}

// out = x - y ; optional copies of out into out2 ; |out|^2 -> dst   (r = chi - A psi ; p = r / r0 = r)
template <typename R>
__global__ void __launch_bounds__(BLAS_BLOCK) xmy_norm_kernel(Cx<R>* out, Cx<R>* out2, const Cx<R>* x, const Cx<R>* y, size_t n,
                                                             ReduceBuf red0, double* dst, size_t fstride) {
  const ReduceBuf red = red0.template for_rhs<1>(blockIdx.y);
  if (out) out += blockIdx.y * fstride;
  if (out2) out2 += blockIdx.y * fstride;
  x += blockIdx.y * fstride; y += blockIdx.y * fstride; dst += blockIdx.y * S_COUNT;
  double s[1] = {0.0};
  for (size_t i = (size_t)blockIdx.x * BLAS_BLOCK + threadIdx.x; i < n; i += (size_t)gridDim.x * BLAS_BLOCK) {
    const Cx<R> v = csub(x[i], y[i]);
    if (out) out[i] = v;
    if (out2) out2[i] = v;
    s[0] += (double)v.x * v.x + (double)v.y * v.y;
  }
  grid_reduce<1, BLAS_BLOCK>(s, red, FinStore{dst, 1});
}

// out = a*x + b*y (real a,b from the host; setup only).  out may alias x or y (b200_qprop updates in place): no __restrict__.
template <typename R>
__global__ void __launch_bounds__(BLAS_BLOCK) axpby_kernel(Cx<R>* out, double a, const Cx<R>* x, double b,
                                                          const Cx<R>* y, size_t n, size_t fstride) {
  out += blockIdx.y * fstride; x += blockIdx.y * fstride; y += blockIdx.y * fstride;
  for (size_t i = (size_t)blockIdx.x * BLAS_BLOCK + threadIdx.x; i < n; i += (size_t)gridDim.x * BLAS_BLOCK) {
    const Cx<R> xv = x[i], yv = y[i];
    out[i] = mk<R>((R)(a * xv.x + b * yv.x), (R)(a * xv.y + b * yv.y));
  }
}

__global__ void set_scalars_kernel(double* scal, int* status, const double* vals, const int* slots, int n, int reset_status);

}  // namespace b200
