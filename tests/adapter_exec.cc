// adapter_exec.cc -- RUNS the Chroma adapter (chroma_adapter/*.{h,cc}) end to end against the real libb200clover.so:
//   TheLinOpFermSystemSolverFactory["B200_CLOVER_INVERTER"] -> createFerm(xml, path, state, A) -> constructor (b200_create,
//   b200_load_gauge from state->getLinks(), b200_make_clover, checkOperator(A)) -> operator()(psi, chi) -> residual re-check
//   with the caller's own A -- the call sequence of quarkprop4_w.cc:280-293 / syssolver_linop_clover_quda_w.h:71-648 --
// and the same for the MdagM and multi-shift plugins.  QDP++ is replaced by tests/mock_chroma (functional stand-in); the
// caller's linear operator A is the CPU oracle (liboracle.so: orc_op_apply), i.e. independent of the engine.
// With nranks = 2 the process forks one rank per T slab (both on CUDA device 0), the mock's QDPInternal::globalSumArray
// runs over a shared-memory page, and B200Glue::allgather / barrier carry the CUDA IPC bootstrap of the engine.
// TEST INFRASTRUCTURE (tests/test_adapter_exec.py builds and runs it; -m gpu).
//
// usage: adapter_exec Lx Ly Lz Lt nranks
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cstdio>
#include <cstring>
#include <stdexcept>

#include "chromabase.h"
#include "actions/ferm/invert/syssolver_linop_factory.h"
#include "actions/ferm/invert/syssolver_mdagm_factory.h"
#include "actions/ferm/invert/multi_syssolver_mdagm_factory.h"
#include "actions/ferm/invert/b200_solvers/syssolver_linop_clover_b200_w.h"
#include "actions/ferm/invert/b200_solvers/syssolver_mdagm_clover_b200_w.h"
#include "actions/ferm/invert/b200_solvers/multi_syssolver_mdagm_clover_b200_w.h"

extern "C" {   // oracle/oracle.c
void* orc_op_create(const int L[4], const double* const u[4], double Mass, double clovCoeffR, double clovCoeffT, int anisoP, int t_dir,
                    double xi_0, double nu);
void orc_op_apply(void* op, double* chi, const double* psi, int isign);
void orc_op_set_symmetric(void* op, int sym);
void orc_op_set_twisted_mass(void* op, double mu);
}
namespace QDP { namespace QDPInternal { void mockAttachShared(void* page); } }

using namespace Chroma;

namespace {
typedef LatticeFermion T;
typedef multi1d<LatticeColorMatrix> Q;

struct SharedPage { char sync[sizeof(int) * (2 + (1 << 16))]; double field[1 << 20]; };
SharedPage* g_page = 0;

int g_rank = 0, g_nranks = 1, g_L[4], g_Lloc[4];

int cb2(const int L[4], const int c[4]) {
  const int V = L[0] * L[1] * L[2] * L[3];
  return ((c[0] + c[1] + c[2] + c[3]) & 1) * (V / 2) + ((c[3] * L[2] + c[2]) * L[1] + c[1]) * (L[0] / 2) + c[0] / 2;
}
// global cb2 index of every local site of this rank's T slab
std::vector<int> local_to_global() {
  std::vector<int> m(g_Lloc[0] * g_Lloc[1] * g_Lloc[2] * g_Lloc[3]);
  int c[4];
  for (c[3] = 0; c[3] < g_Lloc[3]; ++c[3]) for (c[2] = 0; c[2] < g_Lloc[2]; ++c[2]) for (c[1] = 0; c[1] < g_Lloc[1]; ++c[1]) for (c[0] = 0; c[0] < g_Lloc[0]; ++c[0]) {
    int gc[4] = {c[0], c[1], c[2], c[3] + g_rank * g_Lloc[3]};
    m[cb2(g_Lloc, c)] = cb2(g_L, gc);
  }
  return m;
}

// the fermion action's linear operator as the plugin sees it: the CPU oracle on the GLOBAL lattice.  On two ranks the
// local pieces of the argument are assembled through the shared page and every rank applies the global operator.
class OracleLinOp : public LinearOperator<T> {
 public:
  OracleLinOp(void* op_, const std::vector<int>& map_) : op(op_), map(map_), Vg(g_L[0] * g_L[1] * g_L[2] * g_L[3]), in(24 * Vg), out(24 * Vg) {}
  void operator()(T& chi, const T& psi, enum PlusMinus isign) const {
    const int Vl = Layout::sitesOnNode();
    double* buf = g_nranks > 1 ? g_page->field : in.data();
    if (g_nranks > 1) { int one = 1; QDPInternal::globalSum(one); }           // everybody is done reading the previous buffer
    for (int s = Vl / 2; s < Vl; ++s) std::memcpy(buf + 24 * (size_t)map[s], psi.words() + 24 * (size_t)s, 24 * sizeof(double));
    if (g_nranks > 1) { int one = 1; QDPInternal::globalSum(one); std::memcpy(in.data(), buf, sizeof(double) * 24 * Vg); }
    orc_op_apply(op, out.data(), in.data(), isign == PLUS ? +1 : -1);
    for (int s = Vl / 2; s < Vl; ++s) std::memcpy(chi.words() + 24 * (size_t)s, out.data() + 24 * (size_t)map[s], 24 * sizeof(double));
  }
  const Subset& subset() const { return rb[1]; }
 private:
  void* op; std::vector<int> map; int Vg; mutable std::vector<double> in, out;
};

class State : public FermState<T, Q, Q> {
 public:
  explicit State(const Q& u_) : u(u_) {}
  const Q& getLinks() const { return u; }
 private:
  Q u;
};

// smooth SU(3) links on the GLOBAL lattice, identical on every rank: Gram-Schmidt of 1 + eps * Gaussian
void make_gauge(std::vector<std::vector<double> >& ug) {
  const int Vg = g_L[0] * g_L[1] * g_L[2] * g_L[3];
  Seed keep; RNG::savern(keep);
  RNG::setrn(Seed(0x1234567ull));
  ug.assign(4, std::vector<double>((size_t)Vg * 18));
  for (int mu = 0; mu < 4; ++mu)
    for (int s = 0; s < Vg; ++s) {
      std::complex<double> m[3][3];
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m[i][j] = std::complex<double>((i == j) + 0.25 * RNG::gauss(), 0.25 * RNG::gauss());
      for (int i = 0; i < 2; ++i) {
        for (int k = 0; k < i; ++k) {
          std::complex<double> d = 0;
          for (int j = 0; j < 3; ++j) d += std::conj(m[k][j]) * m[i][j];
          for (int j = 0; j < 3; ++j) m[i][j] -= d * m[k][j];
        }
        double n = 0;
        for (int j = 0; j < 3; ++j) n += std::norm(m[i][j]);
        for (int j = 0; j < 3; ++j) m[i][j] /= std::sqrt(n);
      }
      for (int j = 0; j < 3; ++j) m[2][j] = std::conj(m[0][(j + 1) % 3] * m[1][(j + 2) % 3] - m[0][(j + 2) % 3] * m[1][(j + 1) % 3]);
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { ug[mu][(size_t)s * 18 + 6 * i + 2 * j] = m[i][j].real(); ug[mu][(size_t)s * 18 + 6 * i + 2 * j + 1] = m[i][j].imag(); }
    }
  // antiperiodic T: U_t *= -1 on the last global time slice (SimpleFermBC, simple_fermbc.h:87-103)
  int c[4];
  for (c[3] = g_L[3] - 1, c[2] = 0; c[2] < g_L[2]; ++c[2]) for (c[1] = 0; c[1] < g_L[1]; ++c[1]) for (c[0] = 0; c[0] < g_L[0]; ++c[0]) {
    const size_t s = cb2(g_L, c);
    for (int k = 0; k < 18; ++k) ug[3][s * 18 + k] = -ug[3][s * 18 + k];
  }
  RNG::setrn(keep);
}

#define CHECK(cond, what) do { if (!(cond)) { std::printf("[rank %d] FAIL: %s\n", g_rank, what); return 1; } else std::printf("[rank %d] ok: %s\n", g_rank, what); } while (0)

// |chi - A psi| / |chi| over this rank's sites (every rank checks its own part)
double rel_resid(const LinearOperator<T>& A, const T& psi, const T& chi) {
  T tmp = zero, r = zero;
  A(tmp, psi, PLUS);
  r[rb[1]] = chi;
  r[rb[1]] -= tmp;
  return std::sqrt(toDouble(norm2(r, rb[1])) / toDouble(norm2(chi, rb[1])));
}

int run_rank() {
  const int grid[4] = {1, 1, 1, g_nranks}, coord[4] = {0, 0, 0, g_rank};
  Layout::mockSetup(g_Lloc, grid, coord, g_rank, g_nranks);
  if (g_nranks > 1) QDPInternal::mockAttachShared(g_page);
  const std::vector<int> map = local_to_global();
  std::vector<std::vector<double> > ug;
  make_gauge(ug);
  Q links(4);
  const int Vl = Layout::sitesOnNode();
  for (int mu = 0; mu < 4; ++mu)
    for (int s = 0; s < Vl; ++s) std::memcpy(links[mu].words() + 18 * (size_t)s, &ug[mu][(size_t)map[s] * 18], 18 * sizeof(double));
  const double* up[4] = {ug[0].data(), ug[1].data(), ug[2].data(), ug[3].data()};
  const double Mass = 0.1, csw = 1.0;
  void* oracle_op = orc_op_create(g_L, up, Mass, csw, csw, 0, 3, 1.0, 1.0);
  Handle< LinearOperator<T> > A(new OracleLinOp(oracle_op, map));
  Handle< FermState<T, Q, Q> > state(new State(links));

  std::map<std::string, std::string> kv;
  kv["InvertParam/invType"] = "B200_CLOVER_INVERTER";
  kv["InvertParam/MaxIter"] = "500";
  kv["InvertParam/RsdTarget"] = "1.0e-9";
  kv["InvertParam/CloverParams/Mass"] = "0.1";
  kv["InvertParam/CloverParams/clovCoeff"] = "1.0";
  kv["InvertParam/AntiPeriodicT"] = "true";
  kv["InvertParam/SolverType"] = "CG";
  kv["InvertParam/Device"] = "0";
  XMLReader xml(kv);

  CHECK(LinOpSysSolverB200CloverEnv::registerAll(), "LinOp plugin registers in TheLinOpFermSystemSolverFactory");
  CHECK(MdagMSysSolverB200CloverEnv::registerAll(), "MdagM plugin registers in TheMdagMFermSystemSolverFactory");
  CHECK(MdagMMultiSysSolverB200CloverEnv::registerAll(), "multi-shift plugin registers in TheMdagMFermMultiSystemSolverFactory");

  // ---- LinOp plugin, CG then BiCGStab
  T chi = zero, psi = zero;
  gaussian(chi, rb[1]);
  for (int pass = 0; pass < 2; ++pass) {
    xml.kv["InvertParam/SolverType"] = pass ? "BICGSTAB" : "CG";
    Seed before, after;
    RNG::savern(before);
    LinOpSystemSolver<T>* solver = TheLinOpFermSystemSolverFactory::Instance().createObject("B200_CLOVER_INVERTER", xml, "/InvertParam", state, A);
    RNG::savern(after);
    CHECK(before == after, "constructing the plugin (checkOperator included) leaves QDP++'s RNG state untouched");
    psi = zero;
    SystemSolverResults_t res = (*solver)(psi, chi);
    const double rr = rel_resid(*A, psi, chi);
    std::printf("[rank %d] %s: n_count %d resid %.3e, |chi - A psi|/|chi| with the oracle operator %.3e\n", g_rank, pass ? "BICGSTAB" : "CG", res.n_count,
                toDouble(res.resid), rr);
    CHECK(res.n_count > 3 && res.n_count < 500, "LinOp plugin operator() iterates and converges");
    CHECK(rr < 1.0e-8, "solution satisfies the caller's own operator");
    delete solver;
  }

  // ---- a parameter group that does not describe the caller's operator must abort in the constructor
  {
    xml.kv["InvertParam/CloverParams/Mass"] = "0.3";
    bool aborted = false;
    try {
      LinOpSystemSolver<T>* bad = TheLinOpFermSystemSolverFactory::Instance().createObject("B200_CLOVER_INVERTER", xml, "/InvertParam", state, A);
      delete bad;
    } catch (const std::runtime_error&) { aborted = true; }
    xml.kv["InvertParam/CloverParams/Mass"] = "0.1";
    CHECK(aborted, "checkOperator aborts (QDP_abort) when the XML mass differs from the fermion action's");
  }

  // ---- MdagM plugin
  {
    xml.kv["InvertParam/SolverType"] = "CG";
    MdagMSystemSolver<T>* solver = TheMdagMFermSystemSolverFactory::Instance().createObject("B200_CLOVER_INVERTER", xml, "/InvertParam", state, A);
    psi = zero;
    SystemSolverResults_t res = (*solver)(psi, chi);
    T tmp = zero, mm = zero;
    (*A)(tmp, psi, PLUS);
    T tmp2 = zero;
    (*A)(tmp2, tmp, MINUS);
    mm[rb[1]] = chi; mm[rb[1]] -= tmp2;
    const double rr = std::sqrt(toDouble(norm2(mm, rb[1])) / toDouble(norm2(chi, rb[1])));
    std::printf("[rank %d] MdagM CG: n_count %d, local |chi - A^dag A psi|/|chi| %.3e\n", g_rank, res.n_count, rr);
    CHECK(res.n_count > 3 && rr < 1.0e-7, "MdagM plugin solves A^dag A psi = chi for the caller's A");
    delete solver;
  }

  // ---- multi-shift plugin
  {
    MdagMMultiSystemSolver<T>* solver = TheMdagMFermMultiSystemSolverFactory::Instance().createObject("B200_CLOVER_INVERTER", xml, "/InvertParam", state, A);
    multi1d<Real> shifts(3);
    shifts[0] = Real(0.01); shifts[1] = Real(0.2); shifts[2] = Real(1.5);
    multi1d<T> sol(3);
    SystemSolverResults_t res = (*solver)(sol, shifts, chi);     // aborts by itself if any shift misses RsdToleranceFactor * RsdTarget
    std::printf("[rank %d] multi-shift: n_count %d\n", g_rank, res.n_count);
    CHECK(res.n_count > 3, "multi-shift plugin runs and passes its own per-shift residual check with the caller's A");
    delete solver;
  }
  std::printf("[rank %d] ADAPTER_EXEC PASSED\n", g_rank);
  return 0;
}
}  // namespace

int main(int argc, char** argv) {
  if (argc < 6) { std::fprintf(stderr, "usage: %s Lx Ly Lz Lt nranks\n", argv[0]); return 2; }
  for (int i = 0; i < 4; ++i) { g_L[i] = std::atoi(argv[1 + i]); g_Lloc[i] = g_L[i]; }
  g_nranks = std::atoi(argv[5]);
  g_Lloc[3] = g_L[3] / g_nranks;
  if (g_nranks == 1) {
    try { return run_rank(); } catch (const std::exception& e) { std::printf("FAIL: exception %s\n", e.what()); return 1; }
  }
  g_page = static_cast<SharedPage*>(mmap(0, sizeof(SharedPage), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0));
  std::memset(g_page, 0, sizeof(SharedPage));
  std::vector<pid_t> kids;
  for (int r = 0; r < g_nranks; ++r) {
    pid_t p = fork();                       // before any CUDA call: every rank gets its own context
    if (p == 0) {
      g_rank = r;
      int rc = 1;
      try { rc = run_rank(); } catch (const std::exception& e) { std::printf("[rank %d] FAIL: exception %s\n", r, e.what()); }
      std::fflush(stdout);
      _exit(rc);
    }
    kids.push_back(p);
  }
  int bad = 0;
  for (size_t i = 0; i < kids.size(); ++i) { int st = 0; waitpid(kids[i], &st, 0); if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) bad = 1; }
  std::printf("ADAPTER_EXEC %s (%d ranks)\n", bad ? "FAILED" : "PASSED", g_nranks);
  return bad;
}
