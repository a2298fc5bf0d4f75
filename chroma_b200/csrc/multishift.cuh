// multishift.cuh -- device side of the multi-shift CG (MInvCG2_a, lib/actions/ferm/invert/minvcg2.cc:74-373):
// (M^dag M + shift_s) psi_s = chi for all shifts at once, from ONE Krylov sequence.
//
// Per iteration the reference does, on the host, 2 M applies, 2 norms, 1 + n_shift `p` updates, n_shift `psi` updates
// and the z / bs / as recurrences of every shift.  Here:
//   * the two operator applies are the fused kernels of the ordinary CG (EPI_M_NORM: d = |M p0|^2 -> S_A = cp/d = -b;
//     EPI_M_CG: r += b M^dag M p0, c = |r|^2 -> S_B = c/cp = a of the next iteration) -- M^dag M p0 is never written;
//   * ms_scalars_kernel (one warp, lane = shift) advances z, bs, as and the per-shift convergence flags on the device
//     (minvcg2.cc:268-340), so the host never sees a scalar;
//   * ms_update_kernel does psi_s -= bs_s p_s of this iteration AND p_s = z_s r + as_s p_s, p0 = r + a p0 of the next
//     one in a single pass that reads r once for all shifts: (4 n_shift + 3) vectors of traffic instead of the
//     reference's (5 n_shift + 3) in 2 n_shift + 1 separate passes.
#pragma once
#include "common.cuh"
#include "blas.cuh"

namespace b200 {

constexpr int MAX_SHIFT = 32;   // one warp advances all shifts (rational approximations in Chroma's RHMC use <= ~20 poles)

struct MsState {
  double shift[MAX_SHIFT];
  double zprev[MAX_SHIFT], zcur[MAX_SHIFT];   // z[1-iz][s], z[iz][s]
  double bs[MAX_SHIFT], as[MAX_SHIFT];        // bs of this iteration, as of the next p update
  double rsd_sq[MAX_SHIFT];                   // |chi|^2 RsdCG_s^2
  double css[MAX_SHIFT];                      // c z_s^2 at the last test (recurrence residual of shift s)
  int conv[MAX_SHIFT];                        // convsP[s]
  int conv_prev[MAX_SHIFT];                   // convsP[s] as it was when this iteration started
  double a, b;                                // a of the next p update, b of the last residual update
  int n_shift, isz;
};

// iter = 0: the set-up before the loop (minvcg2.cc:211-240); iter >= 1: end of loop iteration `iter` (:268-340).
// On entry scal[S_A] = -b (new), scal[S_B] = a of the next iteration, scal[S_CP] = c = |r|^2 (new).
__global__ void ms_scalars_kernel(MsState* __restrict__ ms, const double* __restrict__ scal, int* __restrict__ status, int iter, int check);   // api.cu

// psi_s -= bs_s p_s (shifts not converged when the iteration started); p_s = z_s r + as_s p_s; p0 = r + a p0.
template <typename R>
__global__ void __launch_bounds__(BLAS_BLOCK) ms_update_kernel(Cx<R>* __restrict__ psi, Cx<R>* __restrict__ p, Cx<R>* __restrict__ p0,
                                                              const Cx<R>* __restrict__ r, size_t n, size_t fstride,
                                                              const MsState* __restrict__ ms, const int* __restrict__ status, int iter, int check) {
  const int stop = check ? status[ST_STOP] : 0;
  if (check && status[ST_BREAKDOWN] != 0) return;
  if (stop != 0 && stop < iter) return;
  __shared__ R s_bs[MAX_SHIFT], s_z[MAX_SHIFT], s_as[MAX_SHIFT];
  __shared__ int s_idx[MAX_SHIFT];
  __shared__ int s_n;
  if (threadIdx.x == 0) {
    int m = 0;
    for (int s = 0; s < ms->n_shift; ++s)
      if (!ms->conv_prev[s]) { s_idx[m] = s; s_bs[m] = (R)ms->bs[s]; s_z[m] = (R)ms->zcur[s]; s_as[m] = (R)ms->as[s]; ++m; }
    s_n = m;
  }
  __syncthreads();
  const int m = s_n;
  const R a = (R)ms->a;
  for (size_t i = (size_t)blockIdx.x * BLAS_BLOCK + threadIdx.x; i < n; i += (size_t)gridDim.x * BLAS_BLOCK) {
    const Cx<R> rv = r[i];
    Cx<R> q = p0[i];
    q.x = rv.x + a * q.x; q.y = rv.y + a * q.y;
    p0[i] = q;
    for (int j = 0; j < m; ++j) {
      const size_t o = (size_t)s_idx[j] * fstride + i;
      Cx<R> pv = p[o], xv = psi[o];
      xv.x -= s_bs[j] * pv.x; xv.y -= s_bs[j] * pv.y;
      psi[o] = xv;
      pv.x = s_z[j] * rv.x + s_as[j] * pv.x; pv.y = s_z[j] * rv.y + s_as[j] * pv.y;
      p[o] = pv;
    }
  }
}

// out = x - y - sigma z ; |out|^2 -> dst   (true residual of a shifted system, multi_syssolver_mdagm_cg.h:86-97)
template <typename R>
__global__ void __launch_bounds__(BLAS_BLOCK) shifted_resid_kernel(const Cx<R>* __restrict__ x, const Cx<R>* __restrict__ y, const Cx<R>* __restrict__ z,
                                                                  double sigma, size_t n, ReduceBuf red, double* dst) {
  double s[1] = {0.0};
  for (size_t i = (size_t)blockIdx.x * BLAS_BLOCK + threadIdx.x; i < n; i += (size_t)gridDim.x * BLAS_BLOCK) {
    const Cx<R> xv = x[i], yv = y[i], zv = z[i];
    const double dx = (double)xv.x - (double)yv.x - sigma * (double)zv.x;
    const double dy = (double)xv.y - (double)yv.y - sigma * (double)zv.y;
    s[0] += dx * dx + dy * dy;
  }
  grid_reduce<1, BLAS_BLOCK>(s, red, FinStore{dst, 1});
}

}  // namespace b200
