/*
 * b200_clover.h -- C ABI of the B200-native even-odd preconditioned Wilson-clover engine.
 *
 * This is the drop-in boundary behind Chroma's solver-plugin surface: the adapter class
 * LinOpSysSolverB200Clover (chroma_adapter/) registered in TheLinOpFermSystemSolverFactory as
 * "B200_CLOVER_INVERTER" binds exactly these entry points, the way the QUDA adapter binds
 * loadGaugeQuda / loadCloverQuda / invertQuda / freeGaugeQuda
 * (lib/actions/ferm/invert/quda_solvers/syssolver_linop_clover_quda_w.h:523,552,573-574 and
 *  syssolver_linop_clover_quda_w.cc:98), and the way the CPU wrappers bind the Dslash C ABI
 * init_sse_su3dslash / sse_su3dslash_wilson / free_sse_su3dslash
 * (other_libs/sse_wilson_dslash/include/sse_dslash.h:13-36).  All citations are relative to
 * the Chroma tree.
 *
 * Conventions
 *  - Plain C: pointers and sizes only, no C++ / torch / QDP++ types, no exceptions.
 *  - Every function returns 0 on success, a B200_ERR_* code otherwise; b200_last_error() gives text.
 *  - "Host" pointers are ordinary host memory in QDP++'s native layout:
 *      site order  cb2: idx = cb*Vh + ((t*Lz+z)*Ly+y)*(Lx/2) + x/2, cb = (x+y+z+t)&1, on the LOCAL
 *                  lattice of this rank (shift_table_scalar.cc:155-214)
 *      fermion     REAL[Vh][spin 4][colour 3][re,im] for ONE checkerboard, i.e. what
 *                  &psi.elem(rb[cb].start()).elem(0).elem(0).real() points at
 *                  (syssolver_linop_clover_quda_w.cc:76,85)
 *      gauge       u[mu] = REAL[V][row 3][col 3][re,im], mu = 0..3, fermion boundary phases already
 *                  multiplied in (state->getLinks(); simple_fermbc.h:87-103), NOT transposed, NOT
 *                  anisotropy-scaled (syssolver_linop_clover_quda_w.h:514-516)
 *      clover      PrimitiveClovTriang<REAL>[V]: diag[2][6] then offd[2][15][re,im], 72 reals/site
 *                  (clover_term_qdp_w.h:19-24)
 *    host_prec says whether REAL is float (B200_SINGLE) or double (B200_DOUBLE).
 *  - "dev" entry points work on device-resident checkerboard fields (b200_field) in the engine's
 *    site-major SoA layout; they exist so a caller can keep vectors on the GPU between calls
 *    (the way QUDA's BUILD_QUDA_DEVIFACE_SPINOR path does, syssolver_linop_clover_quda_w.cc:78-92).
 *  - By default the operator is Chroma's ASYMMETRIC even-odd preconditioned clover operator on the odd
 *    checkerboard, M = A_oo - 1/4 D_oe A_ee^-1 D_eo (eoprec_clover_linop_w.cc:142-187), so that
 *    the adapter's residual check with Chroma's own linop passes.  b200_set_preconditioning switches every
 *    entry point to the SYMMETRIC one, M = 1 - 1/4 A_oo^-1 D_oe A_ee^-1 D_eo (seoprec_clover_linop_w.cc:147-193),
 *    for plugins created with a SymEvenOddPrecCloverLinOp.
 *  - There is no CPU fallback: every entry point fails with B200_ERR_CUDA if no sm_100 device.
 */
#ifndef B200_CLOVER_H
#define B200_CLOVER_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200_ctx b200_ctx;
typedef struct b200_field b200_field; /* one device-resident checkerboard fermion */

enum { B200_SINGLE = 4, B200_DOUBLE = 8 };            /* bytes per real */
enum { B200_RECONS_NONE = 18, B200_RECONS_12 = 12 };   /* reals stored per link; enum_quda_io.h:77-80 */
enum { B200_SOLVER_CG = 0, B200_SOLVER_BICGSTAB = 1 }; /* enum_quda_io.h:21-26 */
enum { B200_PLUS = 1, B200_MINUS = -1 };               /* enum PlusMinus */
enum { B200_PRECOND_ASYMMETRIC = 0, B200_PRECOND_SYMMETRIC = 1 }; /* EvenOddPrecCloverLinOp / SymEvenOddPrecCloverLinOp */
enum { B200_MAX_SHIFTS = 32 };                         /* most shifts one b200_invert_multishift call takes */

enum {
  B200_OK = 0,
  B200_ERR_ARG = 1,       /* bad argument (odd dimension, null pointer, unknown enum ...) */
  B200_ERR_CUDA = 2,      /* CUDA runtime error / no usable device */
  B200_ERR_STATE = 3,     /* call out of order (e.g. solve before gauge/clover are loaded) */
  B200_ERR_BREAKDOWN = 4, /* BiCGStab breakdown (rho = 0, <r0|v> = 0, |t| = 0); invbicgstab.cc:80-83,109-112,133-136 */
  B200_ERR_COMM = 5       /* multi-GPU bootstrap failed, or a peer wait (halo flag / reduction mailbox) ran out of its budget */
};

/* Host-side collectives the caller lends the engine for the ONE-TIME multi-GPU bootstrap (exchange of
 * CUDA IPC handles).  In Chroma these are two QMP/MPI calls; in the Python harness torch.distributed.
 * The hot path never calls them: halos and global sums travel GPU-to-GPU over NVLink peer memory. */
typedef struct {
  int rank, size;
  /* gather `bytes` bytes from every rank into recv (size*bytes), rank order */
  int (*allgather)(void* user, const void* send, void* recv, size_t bytes);
  int (*barrier)(void* user);
  void* user;
} b200_comm;

typedef struct {
  int n_count;          /* iterations, as SystemSolverResults_t::n_count (lib/syssolver.h:16-23) */
  int converged;        /* 1 if the recurrence residual met rsd_target */
  double resid;         /* |chi - M psi| recomputed with M, as syssolver_linop_cg.h:80-87 */
  double rel_resid;     /* resid / |chi| */
  double rsd_sq_iter;   /* last recurrence |r|^2 (invcg2.cc:182 / invbicgstab.cc:160) */
  double secs;          /* device time of the iteration loop (CUDA events) */
  double secs_total;    /* including host<->device copies of chi / psi when host pointers were given */
  double gflops;        /* Chroma's flop count / secs: CG 2*M + 240, BiCGStab 2*M + 960 per site-iter */
  int n_updates;        /* b200_invert_reliable: number of fp64 residual replacements (0 for the other solvers) */
  int reserved;
} b200_solve_info;

/* Environment knobs read at b200_create: B200_COPY_THREADS (host threads of the pageable-buffer bounce pipeline; default every
 * core the rank may use, at most 16; 0 = plain cudaMemcpy), B200_PIN_KB (size of its pinned bounce buffers, default 32768),
 * B200_NT_COPY (0: downloads leave the bounce buffer with plain instead of non-temporal stores), B200_PEER_TIMEOUT_S (budget of
 * every multi-GPU peer wait, default 60 s; when it runs out the call returns B200_ERR_COMM), B200_MRHS_L2_KB (L2 budget of the
 * z-chunked traversal), B200_QPROP_BATCH (cap on the right-hand sides b200_qprop solves at once), B200_SPLIT_MIN_BLOCKS (grid size
 * above which the Dslash reductions are finished by a one-CTA kernel; tests). */
const char* b200_last_error(void);
const char* b200_version(void);
int b200_device_count(void);   /* CUDA devices visible to this process (<= 0: none); for rank -> device mapping */

/* Create an engine context on CUDA device `device` for a local lattice.  global_dims/proc_grid/
 * proc_coord describe the 4-D domain decomposition the way Layout::lattSize()/logicalSize()/
 * nodeCoord do; comm may be NULL when the grid is 1x1x1x1.  The grid may split T and Z (1 x 1 x Pz x Pt, at most 8
 * ranks: T slabs first, then T x Z as for BASELINE config 5); local extents must be even.  comm->rank must be
 * pt*Pz + pz for grid coordinate (pz, pt) and comm->size = Pz*Pt.
 * prec = B200_DOUBLE or B200_SINGLE chooses the device storage + arithmetic precision (reductions
 * are always accumulated in double).  Replaces initQuda (lib/init/chroma_init.cc:250). */
int b200_create(b200_ctx** ctx, int device, const int global_dims[4], const int proc_grid[4],
                const int proc_coord[4], const b200_comm* comm, int prec);
void b200_destroy(b200_ctx* ctx); /* freeGaugeQuda/freeCloverQuda/endQuda, syssolver_linop_clover_quda_w.h:570-575 */

/* Upload the four link fields.  aniso_coeff[mu] multiplies U_mu inside the hopping term only
 * (makeFermCoeffs, lib/io/aniso_io.cc:63-80; lwldslash_w_cppd.cc:115-121); pass {1,1,1,1} (or NULL) if isotropic.
 * t_boundary (+1 / -1) says which phase the fermion BC multiplied into U_t on the last global time slice; it is
 * only used with B200_RECONS_12, where the phase is stripped on upload and re-applied in the kernel, like the
 * QUDA / QPhiX adapters do (syssolver_linop_clover_quda_w.h:147-156, syssolver_linop_clover_qphix_w.h:107-116).
 * Replaces loadGaugeQuda (syssolver_linop_clover_quda_w.h:523) and CPPWilsonDslashD::create +
 * qdp_pack_gauge (lwldslash_w_cppd.cc:96-149, qdp_packer_nopad.cc:7-21). */
int b200_load_gauge(b200_ctx* ctx, const void* const u[4], int host_prec, const double aniso_coeff[4],
                    int t_boundary, int reconstruct);

/* Either hand over Chroma's clover term and its cb-0 inverse (both PrimitiveClovTriang[V], as
 * clov.getTriBuffer()/invclov.getTriBuffer() give them) -- replaces loadCloverQuda
 * (syssolver_linop_clover_quda_w.h:552) ... */
int b200_load_clover(b200_ctx* ctx, const void* clov_tri, const void* invclov_tri, int host_prec);
/* ... or let the GPU build both from the loaded links: field strength (lib/meas/glue/mesfield.cc:44-74),
 * makeClov (clover_term_qdp_w.h:398-553) and the per-site LDL^dagger inverse on cb 0 (:619-846).
 * diag_mass, clov_r, clov_t are the values QDPCloverTermT::create derives (:263-278):
 * diag_mass = 1 + 3*(nu/xi_0 | 1) + Mass, clov_r = clovCoeffR/2 (/xi_0 if aniso), clov_t = clovCoeffT/2.
 * Needs the UN-scaled links, so call b200_load_gauge first (it keeps what this needs). */
int b200_make_clover(b200_ctx* ctx, double diag_mass, double clov_r, double clov_t, int aniso, int t_dir);
/* Read back what b200_make_clover / b200_load_clover hold (PrimitiveClovTriang<double>[V]); for tests. */
int b200_get_clover(b200_ctx* ctx, void* clov_tri, void* invclov_tri, int host_prec);
/* sum over cb-0 sites of log|det A_ee| from the LDL^dagger pass (tr_log_diag, clover_term_qdp_w.h:737) */
int b200_clover_logdet(b200_ctx* ctx, double* tr_log_ee);
/* ... and of log|det A_oo| over the cb-1 sites (symmetric preconditioning only: invclov.choles(1),
 * seoprec_clover_linop_w.cc:31-33; logDetOddOddLinOp of the symmetric log-det operator) */
int b200_clover_logdet_oo(b200_ctx* ctx, double* tr_log_oo);

/* Choose the even-odd preconditioning every later call uses (default B200_PRECOND_ASYMMETRIC):
 *   B200_PRECOND_ASYMMETRIC  M = A_oo - 1/4 D_oe A_ee^-1 D_eo        EvenOddPrecCloverLinOp (eoprec_clover_linop_w.cc:142-187)
 *   B200_PRECOND_SYMMETRIC   M = 1 - 1/4 A_oo^-1 D_oe A_ee^-1 D_eo   SymEvenOddPrecCloverLinOp::operator()
 *                         (seoprec_clover_linop_w.cc:147-193; M^dag = 1 - 1/4 D^dag A_ee^-1 D^dag A_oo^-1)
 * A_oo^-1 is derived on the device from the clover term that is (or later gets) loaded or built -- the LDL^dagger
 * inverse of clover_term_qdp_w.h:619-846 on cb 1, i.e. SymEvenOddPrecCloverLinOp::create's invclov.choles(1)
 * (seoprec_clover_linop_w.cc:31-33).  With the symmetric operator b200_clover_apply accepts inverse = 1 on cb 1,
 * b200_qprop follows SymEvenOddPrecActQprop (seoprec_fermact_qprop.cc:41-100); batched (multi-RHS) fields go through the
 * same multi-RHS kernels as with the asymmetric operator. */
int b200_set_preconditioning(b200_ctx* ctx, int preconditioning);

/* Twisted-mass term of the clover operators (CloverFermActParams::twisted_m, clover_fermact_params_w.cc:90-97): every
 * later application adds chi += mu i gamma_5 psi (PLUS) / chi -= mu i gamma_5 psi (MINUS) on the odd checkerboard, with
 * either preconditioning -- EvenOddPrecCloverLinOp::operator() (eoprec_clover_linop_w.cc:174-184) and
 * SymEvenOddPrecCloverLinOp::operator() (seoprec_clover_linop_w.cc:174-184).  gamma_5 = Gamma(15) = diag(1,1,-1,-1) in
 * Chroma's basis.  mu = 0 (the default) switches the term off.  Single- and multi-RHS kernels, all solvers. */
int b200_set_twisted_mass(b200_ctx* ctx, double mu);

/* Wilson hopping term on one checkerboard: out (parity out_cb) = D in (parity 1-out_cb).
 * Replaces Dslash<REAL>::operator() (cpp_dslash_scalar.h:20-105; cpp_dslash_scalar_64bit.cc:35-65)
 * as called from CPPWilsonDslashD::apply (lwldslash_w_cppd.cc:174-215). */
int b200_dslash(b200_ctx* ctx, void* out_cb_host, const void* in_cb_host, int host_prec, int isign, int out_cb);
/* Clover term (inverse = 0) or its inverse (inverse = 1; cb 0 only, unless symmetric preconditioning is on) on checkerboard cb:
 * QDPCloverTermT::apply (clover_term_qdp_w.h:2138-2160). */
int b200_clover_apply(b200_ctx* ctx, void* out_cb_host, const void* in_cb_host, int host_prec, int cb, int inverse);
/* out_odd = M in_odd (isign=+1) or M^dagger in_odd (-1): EvenOddPrecCloverLinOp::operator()
 * (eoprec_clover_linop_w.cc:142-187); also what CloverSchur4D fuses (cpp_clover_scalar_64bit.cc:65-102). */
int b200_clover_matpc(b200_ctx* ctx, void* out_odd_host, const void* in_odd_host, int host_prec, int isign);

/* Solve M psi = chi on the odd checkerboard.  psi holds the initial guess on entry, the solution on exit (an all-zero guess,
 * what quarkprop4_w.cc:74 passes, is recognised by a host scan and set on the device instead of being copied).
 * solver = B200_SOLVER_CG:       LinOpSysSolverCG (syssolver_linop_cg.h:57-96): chi' = M^dag chi, then
 *                                InvCG2_a on M^dag M (invcg2.cc:70-232), stop |r|^2 <= rsd^2 |chi'|^2.
 * solver = B200_SOLVER_BICGSTAB: LinOpSysSolverBiCGStab (syssolver_linop_bicgstab.h:57-95) ->
 *                                InvBiCGStab_a (invbicgstab.cc:10-202), stop |r|^2 < rsd^2 |chi|^2.
 * Replaces invertQuda (syssolver_linop_clover_quda_w.cc:98).  Non-convergence is NOT an error: the call
 * returns 0 with info->converged = 0 and n_count = max_iter, like invcg2.cc:222-228. */
int b200_invert(b200_ctx* ctx, void* psi_odd_host, const void* chi_odd_host, int host_prec, int solver,
                double rsd_target, int max_iter, b200_solve_info* info);

/* Solve M^dag M psi = chi on the odd checkerboard -- the HMC-side shells (TheMdagMFermSystemSolverFactory,
 * syssolver_mdagm_factory.h:30-39; call site two_flavor_monomial_w.h:74-87).
 * solver = B200_SOLVER_CG:       MdagMSysSolverCG::operator() (syssolver_mdagm_cg.h:59-94): InvCG2_a on chi itself.
 * solver = B200_SOLVER_BICGSTAB: MdagMSysSolverBiCGStab::operator() (syssolver_mdagm_bicgstab.h:62-110): Y = M psi,
 *                                solve M^dag Y = chi (InvBiCGStab isign = MINUS), then M psi = Y (PLUS);
 *                                n_count is the sum of both solves.
 * info->resid = |chi - M^dag M psi|.  Replaces invertQuda as called from syssolver_mdagm_clover_quda_w.h. */
int b200_invert_mdagm(b200_ctx* ctx, void* psi_odd_host, const void* chi_odd_host, int host_prec, int solver,
                      double rsd_target, int max_iter, b200_solve_info* info);

/* Mixed-precision reliable-update CG for M psi = chi (normal equations, like B200_SOLVER_CG): fp32 inner iterations,
 * fp64 residual replacement and group-wise solution updates, RelInvCG_a (lib/actions/ferm/invert/reliable_cg.cc:10-190)
 * behind the shell of LinOpSysSolverReliableCGClover (syssolver_linop_rel_cg_clover.h:41-165; XML: RsdTarget, Delta,
 * MaxIter, syssolver_rel_cg_clover_params.cc).  Needs a context created with B200_DOUBLE; the fp32 copy of the gauge
 * and clover fields is made on the device at the first call.  mdagm != 0 solves M^dag M psi = chi instead (no M^dag chi
 * preparation; resid = |chi - M^dag M psi|).  info->n_count counts inner iterations. */
int b200_invert_reliable(b200_ctx* ctx, void* psi_odd_host, const void* chi_odd_host, int host_prec, double rsd_target,
                         double delta, int max_iter, int mdagm, b200_solve_info* info);

/* Mixed-precision reliable-update BiCGStab for M psi = chi: fp32 BiCGStab recurrences, fp64 residual replacement
 * r = b - M x and group-wise solution updates -- RelInvBiCGStab_a (lib/actions/ferm/invert/reliable_bicgstab.cc:13-290)
 * behind LinOpSysSolverReliableBiCGStabClover::operator() (syssolver_linop_rel_bicgstab_clover.h:105-146; XML: RsdTarget,
 * Delta, MaxIter, syssolver_rel_bicgstab_clover_params.cc:15-18).  mdagm != 0: the two-step M^dag M psi = chi solve of
 * MdagMSysSolverReliableBiCGStabClover (syssolver_mdagm_rel_bicgstab_clover.h:104-170): Y = M psi, M^dag Y = chi,
 * M psi = Y; n_count and n_updates are the sums over both solves.  Needs a B200_DOUBLE context. */
int b200_invert_reliable_bicgstab(b200_ctx* ctx, void* psi_odd_host, const void* chi_odd_host, int host_prec,
                                  double rsd_target, double delta, int max_iter, int mdagm, b200_solve_info* info);

/* Multi-shift CG: (M^dag M + shifts[s]) psi[s] = chi for s = 0..n_shift-1 from one Krylov sequence -- MInvCG2_a
 * (lib/actions/ferm/invert/minvcg2.cc:74-373) behind MdagMMultiSysSolverCG::operator()
 * (multi_syssolver_mdagm_cg.h:58-105; factory TheMdagMFermMultiSystemSolverFactory, used by the rational monomials of RHMC).
 * psi_odd_host: n_shift host pointers, every psi[s] is zeroed first like the reference (:118-122); rsd_target: one
 * relative target per shift; the solve stops when every shift's recurrence residual c z_s^2 < rsd_s^2 |chi|^2 (:326-340).
 * info: n_shift entries -- common n_count, and the TRUE residual |chi - (M^dag M + shift_s) psi_s| of each shift
 * (the check the shell logs, multi_syssolver_mdagm_cg.h:84-99).  Not converging within max_iter is reported as
 * info->converged = 0 (the reference aborts, minvcg2.cc:365-367).  1 <= n_shift <= B200_MAX_SHIFTS. */
int b200_invert_multishift(b200_ctx* ctx, void* const psi_odd_host[], const void* chi_odd_host, int host_prec, int n_shift,
                           const double* shifts, const double* rsd_target, int max_iter, b200_solve_info* info);

/* Full-lattice propagator solve for nrhs right-hand sides (the sequential 12 spin-colour loop of
 * quarkprop4_w.cc:70-117 with the even-odd source preparation and solution reconstruction of
 * eoprec_fermact_qprop.cc:41-80 done on the device).  chi/psi: REAL[nrhs][V][4][3][2].  The right-hand sides are solved
 * in batches of up to 12 by the multi-RHS kernels (as many as fit in free device memory; the environment variable
 * B200_QPROP_BATCH caps the batch, 1 = the reference's one-at-a-time loop).  infos: nrhs entries. */
int b200_qprop(b200_ctx* ctx, void* psi_full_host, const void* chi_full_host, int host_prec, int nrhs, int solver,
               double rsd_target, int max_iter, b200_solve_info* infos);

/* ---- device-resident interface ------------------------------------------------------------- */
int b200_field_alloc(b200_ctx* ctx, b200_field** f);
void b200_field_free(b200_ctx* ctx, b200_field* f);
int b200_field_upload(b200_ctx* ctx, b200_field* f, const void* cb_host, int host_prec);
int b200_field_download(b200_ctx* ctx, const b200_field* f, void* cb_host, int host_prec);
int b200_field_zero(b200_ctx* ctx, b200_field* f);
/* Batched fields: nrhs (1..12) checkerboard fermions in one allocation, e.g. the 12 spin-colour sources of a propagator
 * (quarkprop4_w.cc:70-117).  Every b200_dev_* operator accepts them in place of ordinary fields (all arguments of a call
 * must hold the same number of right-hand sides) and then runs the multi-RHS kernels: one CTA serves all right-hand
 * sides of its sites, so links and clover blocks are read from HBM once per batch instead of once per source.  Solvers
 * advance the right-hand sides in lockstep with independent scalars and stopping tests.  Result arrays of
 * b200_dev_norm2 / b200_dev_inner / b200_dev_invert* then hold nrhs entries (b200_dev_inner: nrhs pairs). */
int b200_mfield_alloc(b200_ctx* ctx, int nrhs, b200_field** f);
int b200_mfield_upload(b200_ctx* ctx, b200_field* f, int irhs, const void* cb_host, int host_prec);
int b200_mfield_download(b200_ctx* ctx, const b200_field* f, int irhs, void* cb_host, int host_prec);
int b200_field_nrhs(const b200_field* f);
int b200_dev_dslash(b200_ctx* ctx, b200_field* out, const b200_field* in, int isign, int out_cb);
int b200_dev_clover_apply(b200_ctx* ctx, b200_field* out, const b200_field* in, int cb, int inverse);
int b200_dev_clover_matpc(b200_ctx* ctx, b200_field* out, const b200_field* in, int isign);
int b200_dev_norm2(b200_ctx* ctx, const b200_field* x, double* result);
int b200_dev_inner(b200_ctx* ctx, const b200_field* x, const b200_field* y, double result[2]); /* <x|y> = sum conj(x) y */
int b200_dev_invert(b200_ctx* ctx, b200_field* psi, const b200_field* chi, int solver, double rsd_target,
                    int max_iter, b200_solve_info* info);
int b200_dev_invert_mdagm(b200_ctx* ctx, b200_field* psi, const b200_field* chi, int solver, double rsd_target,
                          int max_iter, b200_solve_info* info);
int b200_dev_invert_reliable(b200_ctx* ctx, b200_field* psi, const b200_field* chi, double rsd_target, double delta,
                             int max_iter, int mdagm, b200_solve_info* info);
int b200_dev_invert_reliable_bicgstab(b200_ctx* ctx, b200_field* psi, const b200_field* chi, double rsd_target, double delta,
                                      int max_iter, int mdagm, b200_solve_info* info);
/* psi: a batched field (b200_mfield_alloc) holding at least n_shift vectors -- up to B200_MAX_SHIFTS; fields with more
 * than 12 vectors are storage only, the batched operators refuse them -- chi: an ordinary field */
int b200_dev_invert_multishift(b200_ctx* ctx, b200_field* psi, const b200_field* chi, int n_shift, const double* shifts,
                               const double* rsd_target, int max_iter, b200_solve_info* info);
/* Benchmark leg: set up the solver recurrences (r, p, ... as the solver's own preamble does), then run exactly
 * n_iter iterations of the loop body per call with the convergence test disabled. */
int b200_dev_iterate_begin(b200_ctx* ctx, b200_field* psi, const b200_field* chi, int solver);
int b200_dev_iterate(b200_ctx* ctx, int solver, int n_iter);

/* Measurement aid: run M (isign) `reps` times and return the average device time in milliseconds of each of its
 * two kernels separately (CUDA events around every launch): ms[0] = fused A_ee^-1 D_eo, ms[1] = fused
 * A_oo x - 1/4 D_oe t.  Used by bench.py for the roofline line. */
int b200_dev_time_matpc(b200_ctx* ctx, b200_field* out, const b200_field* in, int isign, int reps, double ms[2]);

/* Measurement aid for the roofline line of bench.py: with the recurrences of b200_dev_iterate_begin set up, run `reps`
 * iterations of the solver loop with a CUDA event behind every launch and return in ms[0..*n_ms) the average device time
 * (milliseconds) of each launch of one iteration, in launch order -- CG (invcg2.cc:158-220): A_ee^-1 D_eo p | A_oo p -
 * 1/4 D_oe t with |Mp|^2 | A_ee^-1 D_eo^dag Mp | r -= a(..) with |r|^2 | psi += a p, p = r + b p.  BiCGStab
 * (invbicgstab.cc:74-170): p update | A^-1 D | M with <r0|v> | r -= alpha v | A^-1 D | M with <t|r>,|t|^2 | psi, r update. */
int b200_dev_time_solver_kernels(b200_ctx* ctx, int solver, int reps, double* ms, int max_ms, int* n_ms);

/* ---- plumbing -------------------------------------------------------------------------------- */
void* b200_stream(b200_ctx* ctx);              /* cudaStream_t the engine launches on (for event timing) */
int b200_sync(b200_ctx* ctx);
long long b200_launch_count(b200_ctx* ctx);    /* kernels launched by the engine since creation */
int b200_host_alloc(void** p, size_t bytes);   /* pinned host memory (optional, speeds up host<->device copies) */
void b200_host_free(void* p);
int b200_local_volume(const b200_ctx* ctx, int local_dims[4]);

#ifdef __cplusplus
}
#endif
#endif /* B200_CLOVER_H */
