// engine_mixed.cu -- mixed-precision reliable-update CG: fp32 inner recurrence on Engine<float>, fp64 residual
// replacement and group-wise solution updates on Engine<double>.
//
// Follows RelInvCG_a (lib/actions/ferm/invert/reliable_cg.cc:10-190) step by step, behind the shell of
// LinOpSysSolverReliableCGClover (syssolver_linop_rel_cg_clover.h:41-165).  What is B200-specific is WHO decides:
// in the reference the host evaluates updateR / updateX after every iteration (:113-121); here the finaliser of the
// fused |r|^2 reduction (FinRelCp, dslash.cuh) takes the decision on the device and the fp64 kernels of the
// replacement step are PREDICATED launches: they are enqueued every iteration and return at once unless the flag is
// set.  The host never waits for a scalar, so the fp32 iterations stream back to back.
//
//   per iteration k (fp32 unless marked):
//     t  = A_ee^-1 D_eo p ; mp = A_oo p - 1/4 D_oe t ; d = |mp|^2 ; a = c/d                 (EPI_AINV, EPI_M_NORM)
//     t  = A_ee^-1 D_eo^dag mp ; r -= a (A_oo mp - 1/4 D_oe^dag t) ; cp = |r|^2 ; decide     (EPI_AINV, EPI_M_CGREL)
//     x += a p ; if updateR: xd = (double) x  else p = r + b p                               rel_update_kernel
//     [updateR] fp64: rd = b - M^dag M xd ; r = (float) rd ; |rd|^2 ; [updateX] psi += xd, x = 0, b = rd   (4 Dslash + rel_replace_kernel)
//     [updateR] p = r + b p                                                                  rel_p_kernel
//   after the loop: psi += (double) x.
#include <type_traits>

#include "engine_impl.cuh"

namespace b200 {

// ------------------------------------------------------------------------------------------------ conversions
__global__ void __launch_bounds__(BLAS_BLOCK) planes_d2f_kernel(float2* __restrict__ dst, const double2* __restrict__ src, size_t n) {
  for (size_t i = (size_t)blockIdx.x * BLAS_BLOCK + threadIdx.x; i < n; i += (size_t)gridDim.x * BLAS_BLOCK) {
    const double2 v = src[i];
    dst[i] = make_float2((float)v.x, (float)v.y);
  }
}
// dst += (double) src
__global__ void __launch_bounds__(BLAS_BLOCK) planes_add_f2d_kernel(double2* __restrict__ dst, const float2* __restrict__ src, size_t n) {
  for (size_t i = (size_t)blockIdx.x * BLAS_BLOCK + threadIdx.x; i < n; i += (size_t)gridDim.x * BLAS_BLOCK) {
    const float2 v = src[i];
    double2 d = dst[i];
    d.x += (double)v.x; d.y += (double)v.y;
    dst[i] = d;
  }
}

// ------------------------------------------------------------------------------------------------ loop kernels
// x += a p (reliable_cg.cc:100).  Then either hand x to the fp64 side (updateR) or do the next iteration's
// p = r + beta p (:83-86) right away.  On the converging iteration only x is updated.
__global__ void __launch_bounds__(BLAS_BLOCK) rel_update_kernel(float2* __restrict__ x, float2* __restrict__ p, const float2* __restrict__ r,
                                                               double2* __restrict__ xd, size_t n, BlasCtl c) {
  const int stop = c.status[ST_STOP];
  if (c.status[ST_BREAKDOWN] != 0 || (stop != 0 && stop < c.iter)) return;
  const bool conv = (stop == c.iter);
  const bool upd_r = c.status[ST_UPD_R] != 0;
  const float a = (float)c.scal[S_A], b = (float)c.scal[S_B];
  for (size_t i = (size_t)blockIdx.x * BLAS_BLOCK + threadIdx.x; i < n; i += (size_t)gridDim.x * BLAS_BLOCK) {
    float2 pv = p[i], xv = x[i];
    xv.x += a * pv.x; xv.y += a * pv.y;
    x[i] = xv;
    if (upd_r) xd[i] = make_double2((double)xv.x, (double)xv.y);
    else if (!conv) {
      const float2 rv = r[i];
      pv.x = rv.x + b * pv.x; pv.y = rv.y + b * pv.y;
      p[i] = pv;
    }
  }
}

// After rd = b - M^dag M xd: maxrr = |rd|, (updateX: r0Norm = maxrx = |rd|), beta for the next iteration and the
// convergence test on the TRUE residual (reliable_cg.cc:137-163).
struct FinRelReplace {
  double* scal; int* status; int iter;
  __device__ void operator()(const double* t) const {
    const double r_sq = t[0], rnorm = sqrt(r_sq);
    scal[S_MAXRR] = rnorm;
    if (status[ST_UPD_X]) { scal[S_R0NORM] = rnorm; scal[S_MAXRX] = rnorm; }
    const double c = scal[S_C];
    scal[S_CP] = r_sq; scal[S_B] = r_sq / c; scal[S_C] = r_sq;
    scal[S_RNORM] = r_sq;                           // BiCGStab keeps |r|^2 here (rho is NOT recomputed, reliable_bicgstab.cc:205-216)
    status[ST_NUPD] += 1;
    if (status[ST_STOP] == 0 && r_sq < scal[S_RSDSQ]) status[ST_STOP] = iter;
  }
};

// rd = b - mmx ; r = (float) rd ; |rd|^2 ; updateX: psi += xd, x = 0, b = rd  (reliable_cg.cc:124-155)
__global__ void __launch_bounds__(BLAS_BLOCK) rel_replace_kernel(double2* __restrict__ bvec, const double2* __restrict__ mmx, float2* __restrict__ r,
                                                                double2* __restrict__ psi, const double2* __restrict__ xd,
                                                                float2* __restrict__ x, size_t n, BlasCtl c) {
  if (c.status[ST_STOP] != 0 || c.status[ST_BREAKDOWN] != 0 || c.status[ST_UPD_R] == 0) return;
  const bool upd_x = c.status[ST_UPD_X] != 0;
  double red[1] = {0.0};
  for (size_t i = (size_t)blockIdx.x * BLAS_BLOCK + threadIdx.x; i < n; i += (size_t)gridDim.x * BLAS_BLOCK) {
    const double2 bv = bvec[i], m = mmx[i];
    const double2 rd = make_double2(bv.x - m.x, bv.y - m.y);
    r[i] = make_float2((float)rd.x, (float)rd.y);
    red[0] += rd.x * rd.x + rd.y * rd.y;
    if (upd_x) {
      double2 pv = psi[i]; const double2 xv = xd[i];
      pv.x += xv.x; pv.y += xv.y;
      psi[i] = pv;
      x[i] = make_float2(0.f, 0.f);
      bvec[i] = rd;
    }
  }
  grid_reduce<1, BLAS_BLOCK>(red, c.red, FinRelReplace{c.scal, c.status, c.iter});
}

// p = r + beta p with the replaced residual (only on updateR iterations that did not converge)
__global__ void __launch_bounds__(BLAS_BLOCK) rel_p_kernel(float2* __restrict__ p, const float2* __restrict__ r, size_t n, BlasCtl c) {
  if (c.status[ST_STOP] != 0 || c.status[ST_BREAKDOWN] != 0 || c.status[ST_UPD_R] == 0) return;
  const float b = (float)c.scal[S_B];
  for (size_t i = (size_t)blockIdx.x * BLAS_BLOCK + threadIdx.x; i < n; i += (size_t)gridDim.x * BLAS_BLOCK) {
    const float2 rv = r[i]; float2 pv = p[i];
    pv.x = rv.x + b * pv.x; pv.y = rv.y + b * pv.y;
    p[i] = pv;
  }
}

// ------------------------------------------------------------------------------------------------ sloppy engine set-up
// Give the fp32 engine the fp64 engine's stream, scalar and status blocks, and an fp32 copy of its operator.
static int adopt(Engine<float>& lo, Engine<double>& hi) {
  B200_CUDA(cudaSetDevice(hi.cfg.device));
  if (lo.stream != hi.stream) {
    B200_CUDA(cudaStreamSynchronize(lo.stream));
    B200_CUDA(cudaStreamDestroy(lo.stream));
    lo.stream = hi.stream; lo.owns_stream = false;
    lo.halo.stream = hi.stream;
    cudaFree(lo.scal); cudaFree(lo.status);
    lo.scal = hi.scal; lo.status = hi.status; lo.owns_scalars = false;
    lo.halo.status_dev = hi.status;
  }
  const size_t Vh = (size_t)hi.g.Vh;
  const int NG = hi.recon / 2;
  if (!lo.gauge || lo.recon != hi.recon) {
    if (lo.gauge) { cudaFree(lo.gauge); lo.gauge = nullptr; }
    B200_CUDA(cudaMalloc(&lo.gauge, sizeof(float2) * 8 * (size_t)NG * Vh));
  }
  lo.recon = hi.recon; lo.ls = hi.ls; lo.t_boundary_ = hi.t_boundary_;
  for (int mu = 0; mu < 4; ++mu) lo.aniso_[mu] = hi.aniso_[mu];
  int rc = lo.alloc_clover(); if (rc) return rc;
  planes_d2f_kernel<<<hi.blas_grid, BLAS_BLOCK, 0, hi.stream>>>(lo.gauge, hi.gauge, 8 * (size_t)NG * Vh);
  planes_d2f_kernel<<<hi.blas_grid, BLAS_BLOCK, 0, hi.stream>>>(lo.clov, hi.clov, 72 * Vh);
  planes_d2f_kernel<<<hi.blas_grid, BLAS_BLOCK, 0, hi.stream>>>(lo.invclov, hi.invclov, 36 * Vh);
  hi.launches += 3;
  B200_CUDA(cudaGetLastError());
  lo.have_trlog = false;
  // symmetric preconditioning: the fp32 twin takes an fp32 copy of A_oo^-1 as well
  lo.sym = hi.sym;
  lo.twisted_m = hi.twisted_m;
  if (hi.sym) {
    if (!lo.invclov_oo) B200_CUDA(cudaMalloc(&lo.invclov_oo, sizeof(float2) * 36 * Vh));
    planes_d2f_kernel<<<hi.blas_grid, BLAS_BLOCK, 0, hi.stream>>>(lo.invclov_oo, hi.invclov_oo, 36 * Vh);
    hi.launches += 1;
    B200_CUDA(cudaGetLastError());
  }
  lo.operator_epoch = hi.operator_epoch;
  return B200_OK;
}

// the fp32 twin of an fp64 engine: created on first use, re-synchronised whenever the operator changed
static int sloppy_twin(Engine<double>& hi, EngineBase** lo_slot) {
  if (!*lo_slot) {
    Config c = hi.cfg; c.prec = B200_SINGLE;
    EngineBase* e = make_engine_float(c);
    int rc = e->init();
    if (rc) { delete e; return rc; }
    *lo_slot = e;
  }
  Engine<float>& lo = *static_cast<Engine<float>*>(*lo_slot);
  if (lo.stream != hi.stream || lo.operator_epoch != hi.operator_epoch) return adopt(lo, hi);
  return B200_OK;
}

int reliable_solve(EngineBase* hi_b, EngineBase** lo_slot, b200_field* psi_f, const b200_field* chi_f, double rsd, double delta,
                   int max_iter, int mdagm, b200_solve_info* info) {
  if (!hi_b || hi_b->cfg.prec != B200_DOUBLE) { set_error("b200_invert_reliable needs a context created with B200_DOUBLE"); return B200_ERR_ARG; }
  Engine<double>& hi = *static_cast<Engine<double>*>(hi_b);
  B200_CUDA(cudaSetDevice(hi.cfg.device));
  int rc = hi.ready(); if (rc) return rc;
  if (!psi_f || !chi_f || !info || psi_f == chi_f || max_iter < 0 || !(rsd >= 0.0) || !(delta > 0.0)) { set_error("b200_invert_reliable: bad argument"); return B200_ERR_ARG; }
  if (psi_f->nrhs != 1 || chi_f->nrhs != 1) { set_error("b200_invert_reliable: one right-hand side at a time"); return B200_ERR_ARG; }
  rc = sloppy_twin(hi, lo_slot); if (rc) return rc;
  Engine<float>& lo = *static_cast<Engine<float>*>(*lo_slot);
  rc = hi.set_batch(1); if (rc) return rc;
  rc = lo.set_batch(1); if (rc) return rc;
  rc = hi.need_ws(hi.nws(7)); if (rc) return rc;
  rc = lo.need_ws(lo.nws(5)); if (rc) return rc;
  typedef double2 CD; typedef float2 CF;
  CD* psi = (CD*)psi_f->d; const CD* chi = (const CD*)chi_f->d;
  CD *tmp1 = hi.W(1), *tmp2 = hi.W(2), *bvec = hi.W(3), *xd = hi.W(4);
  CF *mp = lo.W(1), *p = lo.W(2), *r = lo.W(3), *x = lo.W(4);
  const size_t n = hi.nelem();
  cudaStream_t st = hi.stream;
  memset(info, 0, sizeof(*info));
  B200_CUDA(cudaEventRecord(hi.ev_t0, st));

  // ---- set-up, reliable_cg.cc:40-73
  const CD* rhs = chi;
  if (!mdagm) { rc = hi.apply_M(hi.W(6), chi, -1, EPI_M, nullptr, nullptr, 0, 0); if (rc) return rc; rhs = hi.W(6); }
  rc = hi.norm2_dev(rhs, S_TMP0); if (rc) return rc;
  rc = hi.apply_M(tmp1, psi, +1, EPI_M, nullptr, nullptr, 0, 0); if (rc) return rc;
  rc = hi.apply_M(tmp2, tmp1, -1, EPI_M, nullptr, nullptr, 0, 0); if (rc) return rc;
  rc = hi.xmy_norm_dev(bvec, nullptr, rhs, tmp2, S_TMP1); if (rc) return rc;            // b = chi - M^dag M psi ; r_sq
  planes_d2f_kernel<<<hi.blas_grid, BLAS_BLOCK, 0, st>>>(r, bvec, n);                     // r = b
  planes_d2f_kernel<<<hi.blas_grid, BLAS_BLOCK, 0, st>>>(p, bvec, n);                     // p = r
  hi.launches += 2;
  B200_CUDA(cudaMemsetAsync(x, 0, sizeof(CF) * n, st));
  rc = hi.fetch_scalars(); if (rc) return rc;
  const double chi_norm = hi.h_scal[S_TMP0], r_sq0 = hi.h_scal[S_TMP1], rsd_sq = rsd * rsd * chi_norm;
  info->rsd_sq_iter = r_sq0;
  int n_count = 0, converged = 0, breakdown = 0, n_upd = 0;
  if (r_sq0 <= rsd_sq) converged = 1;    // already there (the reference would divide by d = 0 for an exact guess)
  else {
    ScalarSet s{}; s.reset_status = 1; s.n = 6;
    const int sl[6] = {S_RSDSQ, S_C, S_R0NORM, S_MAXRX, S_MAXRR, S_DELTA};
    const double rn = sqrt(r_sq0);
    const double vl[6] = {rsd_sq, r_sq0, rn, rn, rn, delta};
    for (int i = 0; i < 6; ++i) { s.slots[i] = sl[i]; s.vals[i] = vl[i]; }
    rc = hi.set_scalars(s); if (rc) return rc;

    int k = 1, slot = 0, prev = -1;
    bool done = false;
    while (k <= max_iter && !done) {
      const int nb = std::min(ITER_BATCH, max_iter - k + 1);
      for (int i = 0; i < nb; ++i) {
        const int it = k + i;
        rc = lo.apply_M(mp, p, +1, EPI_M_NORM, nullptr, nullptr, it, 1); if (rc) return rc;
        rc = lo.apply_M(nullptr, mp, -1, EPI_M_CGREL, r, nullptr, it, 1); if (rc) return rc;
        rel_update_kernel<<<hi.blas_grid, BLAS_BLOCK, 0, st>>>(x, p, r, xd, n, hi.ctl(it, 1));
        rc = hi.launched("rel_update"); if (rc) return rc;
        rc = hi.apply_M(tmp1, xd, +1, EPI_M, nullptr, nullptr, it, 1, ST_UPD_R); if (rc) return rc;
        rc = hi.apply_M(tmp2, tmp1, -1, EPI_M, nullptr, nullptr, it, 1, ST_UPD_R); if (rc) return rc;
        rel_replace_kernel<<<hi.blas_grid, BLAS_BLOCK, 0, st>>>(bvec, tmp2, r, psi, xd, x, n, hi.ctl(it, 1));
        rc = hi.launched("rel_replace"); if (rc) return rc;
        rel_p_kernel<<<hi.blas_grid, BLAS_BLOCK, 0, st>>>(p, r, n, hi.ctl(it, 1));
        rc = hi.launched("rel_p"); if (rc) return rc;
      }
      B200_CUDA(cudaMemcpyAsync(hi.h_status + slot * ST_COUNT, hi.status, sizeof(int) * ST_COUNT, cudaMemcpyDeviceToHost, st));
      B200_CUDA(cudaEventRecord(hi.ev_poll[slot], st));
      if (prev >= 0) {
        B200_CUDA(cudaEventSynchronize(hi.ev_poll[prev]));
        if (hi.h_status[prev * ST_COUNT + ST_STOP] != 0 || hi.h_status[prev * ST_COUNT + ST_BREAKDOWN] != 0) done = true;
      }
      prev = slot; slot ^= 1; k += nb;
    }
    B200_CUDA(cudaStreamSynchronize(st));
    const int* stt = hi.h_status + prev * ST_COUNT;
    breakdown = stt[ST_BREAKDOWN];
    converged = stt[ST_STOP] != 0;
    n_count = converged ? stt[ST_STOP] : max_iter;
    n_upd = stt[ST_NUPD];
    // psi += x (reliable_cg.cc:166-171); x is zero if the last iteration was a group update
    planes_add_f2d_kernel<<<hi.blas_grid, BLAS_BLOCK, 0, st>>>(psi, x, n);
    rc = hi.launched("planes_add_f2d"); if (rc) return rc;
  }
  B200_CUDA(cudaEventRecord(hi.ev_t1, st));
  rc = hi.fetch_scalars(); if (rc) return rc;
  if (n_count > 0) info->rsd_sq_iter = hi.h_scal[S_CP];
  info->n_count = n_count; info->converged = converged; info->n_updates = n_upd;
  rc = hi.true_residual(psi, chi, mdagm, info); if (rc) return rc;
  float ms = 0.f;
  B200_CUDA(cudaEventElapsedTime(&ms, hi.ev_t0, hi.ev_t1));
  info->secs = ms * 1e-3; info->secs_total = info->secs;
  const double gvol = (double)hi.g.Vh * hi.nranks();
  // flop count as RelInvCG_a books it: per iteration 2 M + 20*Nc*Ns, per replacement 2 M + 6*Nc*Ns (reliable_cg.cc:87,108-110,138-139)
  const double flops = (2.0 * 3792.0 + 240.0) * n_count + (2.0 * 3792.0 + 72.0) * n_upd;
  info->gflops = info->secs > 0 ? flops * gvol / info->secs * 1e-9 : 0.0;
  if (breakdown >= 90) return hi.comm_timeout(breakdown);
  return B200_OK;
}

// ================================================================================================ reliable BiCGStab
// RelInvBiCGStab_a (lib/actions/ferm/invert/reliable_bicgstab.cc:13-290): the BiCGStab recurrences run in fp32 on the
// sloppy engine (the same five fused launches per iteration as the fp64 solver, engine_impl.cuh::bicg_iteration); the
// finaliser of the fused |r|^2 / <r0|r> reduction (FinBiUpdate with rel = 1) takes the updateR / updateX decisions on
// the device; the fp64 residual replacement r = b - A x (one operator apply + rel_replace_kernel) is enqueued every
// iteration as PREDICATED launches.  rho is not recomputed after a replacement, exactly as in the reference (:205-216).
__global__ void __launch_bounds__(BLAS_BLOCK) planes_f2d_pred_kernel(double2* __restrict__ dst, const float2* __restrict__ src, size_t n,
                                                                    const int* __restrict__ status) {
  if (status[ST_STOP] != 0 || status[ST_BREAKDOWN] != 0 || status[ST_UPD_R] == 0) return;
  for (size_t i = (size_t)blockIdx.x * BLAS_BLOCK + threadIdx.x; i < n; i += (size_t)gridDim.x * BLAS_BLOCK) {
    const float2 v = src[i];
    dst[i] = make_double2((double)v.x, (double)v.y);
  }
}

// One A psi = rhs solve (A = M for isign = +1, M^dag for -1).  hi: W(0) even temporary, W(1) tmp, W(3) b, W(4) x_dble;
// lo: W(1) r, W(2) r0, W(3) p, W(4) v, W(5) t, W(6) x (bicg_iteration's layout).
static int rel_bicg_run(Engine<double>& hi, Engine<float>& lo, double2* psi, const double2* rhs, int isign, double rsd, double delta,
                        int max_iter, int* n_count, int* converged, int* n_upd, double* rsq_iter) {
  typedef double2 CD; typedef float2 CF;
  CD *tmp = hi.W(1), *bvec = hi.W(3), *xd = hi.W(4);
  CF *r = lo.W(1), *r0 = lo.W(2), *p = lo.W(3), *v = lo.W(4), *x = lo.W(6);
  const size_t n = hi.nelem();
  cudaStream_t st = hi.stream;
  // set-up, reliable_bicgstab.cc:56-104
  int rc = hi.norm2_dev(rhs, S_TMP0); if (rc) return rc;
  rc = hi.apply_M(tmp, psi, isign, EPI_M, nullptr, nullptr, 0, 0); if (rc) return rc;
  rc = hi.xmy_norm_dev(bvec, nullptr, rhs, tmp, S_TMP1); if (rc) return rc;              // b = rhs - A psi ; b_sq
  planes_d2f_kernel<<<hi.blas_grid, BLAS_BLOCK, 0, st>>>(r, bvec, n);
  planes_d2f_kernel<<<hi.blas_grid, BLAS_BLOCK, 0, st>>>(r0, bvec, n);
  hi.launches += 2;
  B200_CUDA(cudaMemsetAsync(x, 0, sizeof(CF) * n, st));
  B200_CUDA(cudaMemsetAsync(p, 0, sizeof(CF) * n, st));
  B200_CUDA(cudaMemsetAsync(v, 0, sizeof(CF) * n, st));
  rc = hi.fetch_scalars(); if (rc) return rc;
  const double rhs_sq = hi.h_scal[S_TMP0], r_sq0 = hi.h_scal[S_TMP1], rsd_sq = rsd * rsd * rhs_sq;
  *rsq_iter = r_sq0; *n_count = 0; *converged = 0; *n_upd = 0;
  if (r_sq0 < rsd_sq || r_sq0 == 0.0) { *converged = 1; return B200_OK; }             // nothing to do (rho would be 0)
  {
    ScalarSet s{}; s.reset_status = 1; s.n = 12;
    const double rn = sqrt(r_sq0);
    // rho_1 = |r|^2 (r0 = r), rho_0 = alpha = omega = 1 => beta_1 = rho_1, p_1 = r  (:96-114)
    const int sl[12] = {S_RSDSQ, S_RHO_RE, S_RHO_IM, S_RHOP_RE, S_RHOP_IM, S_ALPHA_RE, S_ALPHA_IM, S_OMEGA_RE, S_OMEGA_IM, S_BETA_RE, S_BETA_IM, S_C};
    const double vl[12] = {rsd_sq, r_sq0, 0.0, 1.0, 0.0, 1.0, 0.0, 1.0, 0.0, r_sq0, 0.0, 1.0};
    for (int i = 0; i < 12; ++i) { s.slots[i] = sl[i]; s.vals[i] = vl[i]; }
    rc = hi.set_scalars(s); if (rc) return rc;
    ScalarSet s2{}; s2.reset_status = 0; s2.n = 4;
    const int sl2[4] = {S_R0NORM, S_MAXRX, S_MAXRR, S_DELTA};
    const double vl2[4] = {rn, rn, rn, delta};
    for (int i = 0; i < 4; ++i) { s2.slots[i] = sl2[i]; s2.vals[i] = vl2[i]; }
    rc = hi.set_scalars(s2); if (rc) return rc;
  }
  lo.ctl_rel = 1;
  int k = 1, slot = 0, prev = -1;
  bool done = false;
  while (k <= max_iter && !done) {
    const int nbat = std::min(ITER_BATCH, max_iter - k + 1);
    for (int i = 0; i < nbat; ++i) {
      const int it = k + i;
      rc = lo.bicg_iteration(x, it, 1, isign); if (rc) { lo.ctl_rel = 0; return rc; }
      planes_f2d_pred_kernel<<<hi.blas_grid, BLAS_BLOCK, 0, st>>>(xd, x, n, hi.status);
      rc = hi.launched("planes_f2d_pred"); if (rc) { lo.ctl_rel = 0; return rc; }
      rc = hi.apply_M(tmp, xd, isign, EPI_M, nullptr, nullptr, it, 1, ST_UPD_R); if (rc) { lo.ctl_rel = 0; return rc; }
      rel_replace_kernel<<<hi.blas_grid, BLAS_BLOCK, 0, st>>>(bvec, tmp, r, psi, xd, x, n, hi.ctl(it, 1));
      rc = hi.launched("rel_replace"); if (rc) { lo.ctl_rel = 0; return rc; }
    }
    B200_CUDA(cudaMemcpyAsync(hi.h_status + slot * ST_COUNT, hi.status, sizeof(int) * ST_COUNT, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaEventRecord(hi.ev_poll[slot], st));
    if (prev >= 0) {
      B200_CUDA(cudaEventSynchronize(hi.ev_poll[prev]));
      if (hi.h_status[prev * ST_COUNT + ST_STOP] != 0 || hi.h_status[prev * ST_COUNT + ST_BREAKDOWN] != 0) done = true;
    }
    prev = slot; slot ^= 1; k += nbat;
  }
  lo.ctl_rel = 0;
  B200_CUDA(cudaStreamSynchronize(st));
  int breakdown = 0;
  if (prev >= 0) {
    const int* stt = hi.h_status + prev * ST_COUNT;
    breakdown = stt[ST_BREAKDOWN];
    *converged = stt[ST_STOP] != 0;
    *n_count = *converged ? stt[ST_STOP] : max_iter;
    *n_upd = stt[ST_NUPD];
  }
  // psi += x (reliable_bicgstab.cc:247-252); x is zero if the last iteration was a group update
  planes_add_f2d_kernel<<<hi.blas_grid, BLAS_BLOCK, 0, st>>>(psi, x, n);
  rc = hi.launched("planes_add_f2d"); if (rc) return rc;
  rc = hi.fetch_scalars(); if (rc) return rc;
  if (*n_count > 0) *rsq_iter = hi.h_scal[S_RNORM];
  if (breakdown >= 90) return hi.comm_timeout(breakdown);
  if (breakdown) { set_error("reliable BiCGStab breakdown (code %d) at iteration <= %d", breakdown, *n_count); return B200_ERR_BREAKDOWN; }
  return B200_OK;
}

// Shells: LinOpSysSolverReliableBiCGStabClover::operator() (syssolver_linop_rel_bicgstab_clover.h:105-146): M psi = chi;
// MdagMSysSolverReliableBiCGStabClover::operator() (syssolver_mdagm_rel_bicgstab_clover.h:104-170): Y = M psi,
// M^dag Y = chi, M psi = Y.
int reliable_bicgstab_solve(EngineBase* hi_b, EngineBase** lo_slot, b200_field* psi_f, const b200_field* chi_f, double rsd, double delta,
                            int max_iter, int mdagm, b200_solve_info* info) {
  if (!hi_b || hi_b->cfg.prec != B200_DOUBLE) { set_error("b200_invert_reliable_bicgstab needs a context created with B200_DOUBLE"); return B200_ERR_ARG; }
  Engine<double>& hi = *static_cast<Engine<double>*>(hi_b);
  B200_CUDA(cudaSetDevice(hi.cfg.device));
  int rc = hi.ready(); if (rc) return rc;
  if (!psi_f || !chi_f || !info || psi_f == chi_f || max_iter < 0 || !(rsd >= 0.0) || !(delta > 0.0)) { set_error("b200_invert_reliable_bicgstab: bad argument"); return B200_ERR_ARG; }
  if (psi_f->nrhs != 1 || chi_f->nrhs != 1) { set_error("b200_invert_reliable_bicgstab: one right-hand side at a time"); return B200_ERR_ARG; }
  rc = sloppy_twin(hi, lo_slot); if (rc) return rc;
  Engine<float>& lo = *static_cast<Engine<float>*>(*lo_slot);
  rc = hi.set_batch(1); if (rc) return rc;
  rc = lo.set_batch(1); if (rc) return rc;
  rc = hi.need_ws(hi.nws(7)); if (rc) return rc;
  rc = lo.need_ws(lo.nws(7)); if (rc) return rc;
  double2* psi = (double2*)psi_f->d; const double2* chi = (const double2*)chi_f->d;
  memset(info, 0, sizeof(*info));
  B200_CUDA(cudaEventRecord(hi.ev_t0, hi.stream));
  int n1 = 0, c1 = 1, u1 = 0, n2 = 0, c2 = 0, u2 = 0;
  double rsq = 0.0;
  if (!mdagm) {
    rc = rel_bicg_run(hi, lo, psi, chi, +1, rsd, delta, max_iter, &n2, &c2, &u2, &rsq);
  } else {
    rc = hi.apply_M(hi.W(6), psi, +1, EPI_M, nullptr, nullptr, 0, 0);
    if (!rc) rc = rel_bicg_run(hi, lo, hi.W(6), chi, -1, rsd, delta, max_iter, &n1, &c1, &u1, &rsq);
    if (!rc) rc = rel_bicg_run(hi, lo, psi, hi.W(6), +1, rsd, delta, max_iter, &n2, &c2, &u2, &rsq);
  }
  info->n_count = n1 + n2; info->converged = c1 && c2; info->n_updates = u1 + u2; info->rsd_sq_iter = rsq;
  if (rc && rc != B200_ERR_BREAKDOWN) return rc;
  B200_CUDA(cudaEventRecord(hi.ev_t1, hi.stream));
  int rc2 = hi.true_residual(psi, chi, mdagm, info); if (rc2) return rc2;
  float ms = 0.f;
  B200_CUDA(cudaEventElapsedTime(&ms, hi.ev_t0, hi.ev_t1));
  info->secs = ms * 1e-3; info->secs_total = info->secs;
  const double gvol = (double)hi.g.Vh * hi.nranks();
  // flops as RelInvBiCGStab_a books them: 2 A + 80*Nc*Ns per iteration, A + 6*Nc*Ns per replacement (:188-190, :222-223)
  const double flops = (2.0 * 3792.0 + 960.0) * info->n_count + (3792.0 + 72.0) * info->n_updates;
  info->gflops = info->secs > 0 ? flops * gvol / info->secs * 1e-9 : 0.0;
  return rc;
}

}  // namespace b200
