// ref_chroma_shim.cc -- extern "C" doors into the REFERENCE'S OWN Chroma-level code for the hot path, compiled unmodified
// from /root/reference against tests/mock_chroma (a functional stand-in for the slice of QDP++ they touch):
//   * the clover site loops of lib/actions/ferm/linop/clover_term_qdp_w.h (makeClovSiteLoop :398-521, LDagDLInvSiteLoop
//     :619-818, applySiteLoop :1562-1634), lifted at build time by oracle/extract_clover_loops.py
//   * mesField, lib/meas/glue/mesfield.cc, whole file
//   * the solver loops lib/actions/ferm/invert/{invcg2,invbicgstab,minvcg2,reliable_cg,reliable_bicgstab}.cc, whole files
// TEST INFRASTRUCTURE: this pins oracle/oracle.c's restatements (tests/test_oracle.py) and writes tests/golden fixtures
// (tests/golden/make_golden.py).  The product never links it.  The linear operator the solvers iterate on is a callback
// (the oracle's own orc_op_apply, itself pinned to the reference Dslash / CloverSchur4D).
#include "chromabase.h"
#include "linearop.h"
#include "syssolver.h"
#include "clover_site_loops.h"
#include "meas/glue/mesfield.h"
#include "actions/ferm/invert/invcg2.h"
#include "actions/ferm/invert/invbicgstab.h"
#include "actions/ferm/invert/minvcg2.h"
#include "actions/ferm/invert/reliable_cg.h"
#include "actions/ferm/invert/reliable_bicgstab.h"

#include <cstring>
#include <vector>

using namespace Chroma;

namespace {
typedef void (*apply_fn)(void* user, double* chi, const double* psi, int isign);   // full-lattice [V][4][3][2] arrays

// LinearOperator<LatticeFermionD> on rb[1] behind a C callback; records |input|^2 of every application
class CallbackLinOp : public LinearOperator<LatticeFermionD> {
 public:
  CallbackLinOp(apply_fn f_, void* user_, double* trace_, int trace_cap_) : f(f_), user(user_), trace(trace_), cap(trace_cap_), n(0) {}
  void operator()(LatticeFermionD& chi, const LatticeFermionD& psi, enum PlusMinus isign) const {
    if (trace && n < cap) trace[n] = toDouble(norm2(psi, rb[1]));
    ++n;
    f(user, chi.words(), psi.words(), isign == PLUS ? +1 : -1);
  }
  const Subset& subset() const { return rb[1]; }
  int calls() const { return n; }
 private:
  apply_fn f; void* user; double* trace; int cap; mutable int n;
};

void to_field(LatticeFermionD& x, const double* p) { std::memcpy(x.words(), p, sizeof(double) * 24 * Layout::sitesOnNode()); }
void from_field(double* p, const LatticeFermionD& x) { std::memcpy(p, x.words(), sizeof(double) * 24 * Layout::sitesOnNode()); }
}  // namespace

extern "C" {

void refc_setup(const int L[4]) {
  const int one[4] = {1, 1, 1, 1}, zero4[4] = {0, 0, 0, 0};
  Layout::mockSetup(L, one, zero4, 0, 1);
}

// mesField (lib/meas/glue/mesfield.cc:30-78) on a single-rank periodic lattice: u = 4 planes of [V][3][3][2], f = 6 planes
void refc_mesfield(const double* u, double* f) {
  const int V = Layout::sitesOnNode();
  multi1d<LatticeColorMatrixD> U(4), F;
  for (int mu = 0; mu < 4; ++mu) std::memcpy(U[mu].words(), u + (size_t)mu * V * 18, sizeof(double) * 18 * V);
  mesField(F, U);
  for (int k = 0; k < 6; ++k) std::memcpy(f + (size_t)k * V * 18, F[k].words(), sizeof(double) * 18 * V);
}

// QDPCloverTermT::makeClov (clover_term_qdp_w.h:525-553): f_k *= getCloverCoeff, then the reference site loop.
// f: 6 planes of [V][3][3][2]; tri: [V][72]
void refc_make_clov(const double* f, const double coef[6], double diag_mass, double* tri) {
  const int V = Layout::sitesOnNode();
  std::vector<LatticeColorMatrixD> F(6);
  for (int k = 0; k < 6; ++k) {
    std::memcpy(F[k].words(), f + (size_t)k * V * 18, sizeof(double) * 18 * V);
    F[k] = Real(coef[k]) * F[k];
  }
  typedef QDPCloverEnv::QDPCloverMakeClovArg<LatticeColorMatrixD> Arg;
  Arg::RealT dm(diag_mass);
  Arg arg = {dm, F[0], F[1], F[2], F[3], F[4], F[5], reinterpret_cast<PrimitiveClovTriang<double>*>(tri)};
  QDPCloverEnv::makeClovSiteLoop<LatticeColorMatrixD>(0, V, 0, &arg);
}

// QDPCloverTermT::ldagdlinv (:821-846): tr_log_diag = zero, then the reference site loop over rb[cb]
void refc_ldagdlinv(double* tri, int cb, double* tr_log) {
  const int V = Layout::sitesOnNode();
  typedef QDPCloverEnv::LDagDLInvArgs<LatticeColorMatrixD> Arg;
  Arg::LatticeRealT tl;
  Arg arg = {tl, reinterpret_cast<PrimitiveClovTriang<double>*>(tri), cb};
  QDPCloverEnv::LDagDLInvSiteLoop<LatticeColorMatrixD>(0, rb[cb].numSiteTable(), 0, &arg);
  for (int i = 0; i < V; ++i) tr_log[i] = tl.elem(i).elem().elem().elem();
}

// QDPCloverTermT::apply (:2138-2160) on checkerboard cb
void refc_clover_apply(const double* tri, const double* psi, double* chi, int cb) {
  LatticeFermionD x, y;
  to_field(x, psi);
  QDPCloverEnv::ApplyArgs<LatticeFermionD> arg = {y, x, reinterpret_cast<const PrimitiveClovTriang<double>*>(tri), cb};
  QDPCloverEnv::applySiteLoop<LatticeFermionD>(0, rb[cb].numSiteTable(), 0, &arg);
  from_field(chi, y);
}

// InvCG2 (invcg2.cc:70-232).  trace[i] = |input|^2 of the i-th operator application (2 in the preamble, 2 per iteration, 2 at
// the end); out = {n_count, resid, number of operator applications}
void refc_invcg2(apply_fn f, void* user, const double* chi, double* psi, double rsd, int maxcg, double* out, double* trace, int trace_cap) {
  LatticeFermionD c, p;
  to_field(c, chi); to_field(p, psi);
  CallbackLinOp M(f, user, trace, trace_cap);
  SystemSolverResults_t r = InvCG2(M, c, p, Real(rsd), maxcg);
  from_field(psi, p);
  out[0] = r.n_count; out[1] = toDouble(r.resid); out[2] = M.calls();
}

// InvBiCGStab (invbicgstab.cc:10-202)
void refc_invbicgstab(apply_fn f, void* user, const double* chi, double* psi, double rsd, int maxit, int isign, double* out, double* trace,
                      int trace_cap) {
  LatticeFermionD c, p;
  to_field(c, chi); to_field(p, psi);
  CallbackLinOp M(f, user, trace, trace_cap);
  SystemSolverResults_t r = InvBiCGStab(M, c, p, Real(rsd), maxit, isign > 0 ? PLUS : MINUS);
  from_field(psi, p);
  out[0] = r.n_count; out[1] = toDouble(r.resid); out[2] = M.calls();
}

// MInvCG2 (minvcg2.cc:74-373): psi = nshift full-lattice fields, back to back
void refc_minvcg2(apply_fn f, void* user, const double* chi, double* psi, const double* shifts, const double* rsd, int nshift, int maxcg,
                  double* out, double* trace, int trace_cap) {
  LatticeFermionD c;
  to_field(c, chi);
  multi1d<LatticeFermionD> p(nshift);
  multi1d<RealD> sh(nshift), rs(nshift);
  for (int s = 0; s < nshift; ++s) { sh[s] = RealD(shifts[s]); rs[s] = RealD(rsd[s]); }
  CallbackLinOp M(f, user, trace, trace_cap);
  int n_count = 0;
  MInvCG2(M, c, p, sh, rs, maxcg, n_count);
  const size_t n = (size_t)24 * Layout::sitesOnNode();
  for (int s = 0; s < nshift; ++s) from_field(psi + s * n, p[s]);
  out[0] = n_count; out[1] = 0; out[2] = M.calls();
}

// InvCGReliable with an fp32 inner operator and an fp64 outer one (reliable_cg.cc:10-190): both callbacks get double arrays,
// the fp32 one works on copies rounded to float (its input and output pass through float fields)
struct F32Ctx { apply_fn f; void* user; };
}  // extern "C"

namespace {
class CallbackLinOpF : public LinearOperator<LatticeFermionF> {
 public:
  CallbackLinOpF(apply_fn f_, void* user_) : f(f_), user(user_), n(0) {}
  void operator()(LatticeFermionF& chi, const LatticeFermionF& psi, enum PlusMinus isign) const {
    const int nw = 24 * Layout::sitesOnNode();
    std::vector<double> in(nw), out(nw);
    for (int i = 0; i < nw; ++i) in[i] = psi.words()[i];
    f(user, out.data(), in.data(), isign == PLUS ? +1 : -1);
    for (int i = 0; i < nw; ++i) chi.words()[i] = (float)out[i];
    ++n;
  }
  const Subset& subset() const { return rb[1]; }
  int calls() const { return n; }
 private:
  apply_fn f; void* user; mutable int n;
};
}  // namespace

extern "C" void refc_reliable_cg(apply_fn f, void* user, const double* chi, double* psi, double rsd, double delta, int maxit, double* out) {
  LatticeFermionD c, p;
  to_field(c, chi); to_field(p, psi);
  CallbackLinOp M(f, user, 0, 0);
  CallbackLinOpF Mf(f, user);
  SystemSolverResults_t r = InvCGReliable(M, Mf, c, p, Real(rsd), Real(delta), maxit);
  from_field(psi, p);
  out[0] = r.n_count; out[1] = toDouble(r.resid); out[2] = M.calls(); out[3] = Mf.calls();
}

// InvBiCGStabReliable with an fp32 inner operator (reliable_bicgstab.cc:13-341)
extern "C" void refc_reliable_bicgstab(apply_fn f, void* user, const double* chi, double* psi, double rsd, double delta, int maxit, int isign,
                                       double* out) {
  LatticeFermionD c, p;
  to_field(c, chi); to_field(p, psi);
  CallbackLinOp M(f, user, 0, 0);
  CallbackLinOpF Mf(f, user);
  SystemSolverResults_t r = InvBiCGStabReliable(M, Mf, c, p, Real(rsd), Real(delta), maxit, isign > 0 ? PLUS : MINUS);
  from_field(psi, p);
  out[0] = r.n_count; out[1] = toDouble(r.resid); out[2] = M.calls(); out[3] = Mf.calls();
}
