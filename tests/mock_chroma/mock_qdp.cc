// Definitions behind tests/mock_chroma/chromabase.h (TEST INFRASTRUCTURE): layout, subsets, RNG, XML key/value store,
// and cross-rank sums for harnesses that fork one process per rank (a shared-memory page + a sense-reversing barrier).
#include "chromabase.h"

#include <algorithm>
#include <sstream>
#include <stdexcept>

namespace QDP {
Subset all;
Subset rb[2];
Zero zero;

namespace {
multi1d<int> g_latt(4), g_sub(4), g_grid(4), g_coord(4);
int g_node = 0, g_nodes = 1, g_sites = 0;
uint64_t g_rng = 0x9E3779B97F4A7C15ull;
struct Shared { volatile int count; volatile int sense; int buf[1 << 16]; };
Shared* g_shared = 0;
}

namespace Layout {
void mockSetup(const int local_dims[4], const int grid[4], const int coord[4], int node, int nodes) {
  g_sites = 1;
  for (int i = 0; i < 4; ++i) {
    g_sub[i] = local_dims[i]; g_grid[i] = grid[i]; g_coord[i] = coord[i];
    g_latt[i] = local_dims[i] * grid[i];
    g_sites *= local_dims[i];
  }
  g_node = node; g_nodes = nodes;
  all.make(0, g_sites);
  rb[0].make(0, g_sites / 2);          // cb2 order: checkerboard 0 first (shift_table_scalar.cc:190-214)
  rb[1].make(g_sites / 2, g_sites);
}
// cb2 site index of local coordinate c (shift_table_scalar.cc:190-214)
static int cb2_index(const int c[4]) {
  const int* L = g_sub.slice();
  const int cb = (c[0] + c[1] + c[2] + c[3]) & 1;
  return cb * (g_sites / 2) + ((c[3] * L[2] + c[2]) * L[1] + c[1]) * (L[0] / 2) + c[0] / 2;
}
const int* mockNeighbourTable(int dir, int mu) {
  static std::vector<int> tab[2][4];
  static int built_for = -1;
  if (built_for != g_sites) { for (int d = 0; d < 2; ++d) for (int m = 0; m < 4; ++m) tab[d][m].clear(); built_for = g_sites; }
  std::vector<int>& t = tab[dir > 0 ? 1 : 0][mu];
  if (t.empty()) {
    const int* L = g_sub.slice();
    t.resize(g_sites);
    int c[4];
    for (c[3] = 0; c[3] < L[3]; ++c[3]) for (c[2] = 0; c[2] < L[2]; ++c[2]) for (c[1] = 0; c[1] < L[1]; ++c[1]) for (c[0] = 0; c[0] < L[0]; ++c[0]) {
      int n[4] = {c[0], c[1], c[2], c[3]};
      n[mu] = (c[mu] + (dir > 0 ? 1 : L[mu] - 1)) % L[mu];
      t[cb2_index(c)] = cb2_index(n);
    }
  }
  return t.data();
}
const multi1d<int>& lattSize() { return g_latt; }
const multi1d<int>& subgridLattSize() { return g_sub; }
const multi1d<int>& logicalSize() { return g_grid; }
const multi1d<int>& nodeCoord() { return g_coord; }
int nodeNumber() { return g_node; }
int numNodes() { return g_nodes; }
int sitesOnNode() { return g_sites; }
int vol() { return g_sites * g_nodes; }
}  // namespace Layout

namespace RNG {
void savern(Seed& s) { s = g_rng; }
void setrn(const Seed& s) { g_rng = s; }
static double uniform() {
  g_rng = g_rng * 6364136223846793005ull + 1442695040888963407ull;
  return ((g_rng >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}
double gauss() { return std::sqrt(-2.0 * std::log(uniform())) * std::cos(6.283185307179586 * uniform()); }
}  // namespace RNG

namespace QDPInternal {
void mockAttachShared(void* page) { g_shared = static_cast<Shared*>(page); }
static void barrier_() {
  if (!g_shared || g_nodes == 1) return;
  const int my = 1 - g_shared->sense;
  if (__sync_add_and_fetch(&g_shared->count, 1) == g_nodes) { g_shared->count = 0; __sync_synchronize(); g_shared->sense = my; }
  else while (g_shared->sense != my) { __sync_synchronize(); }
}
void globalSumArray(int* a, int n) {
  if (!g_shared || g_nodes == 1) return;
  if (n > (1 << 16)) throw std::runtime_error("mock globalSumArray: buffer too small");
  barrier_();
  if (g_node == 0) for (int i = 0; i < n; ++i) g_shared->buf[i] = 0;
  barrier_();
  for (int i = 0; i < n; ++i) if (a[i]) __sync_fetch_and_add(&g_shared->buf[i], a[i]);
  barrier_();
  for (int i = 0; i < n; ++i) a[i] = g_shared->buf[i];
  barrier_();
}
void globalSum(int& x) { globalSumArray(&x, 1); }
}  // namespace QDPInternal

namespace QDPIO { std::ostream& cout = std::cout; std::ostream& cerr = std::cerr; }
void QDP_abort(int rc) { std::ostringstream m; m << "QDP_abort(" << rc << ")"; throw std::runtime_error(m.str()); }

void QDP_error_exit(const char* fmt, ...) { QDPIO::cerr << "QDP_error_exit: " << fmt << std::endl; QDP_abort(1); }

std::string XMLReader::get(const std::string& tag) const {
  std::map<std::string, std::string>::const_iterator it = kv.find(join(prefix, tag));
  if (it == kv.end()) { QDPIO::cerr << "XMLReader: no tag " << join(prefix, tag) << std::endl; QDP_abort(1); }
  return it->second;
}
void read(XMLReader& x, const std::string& t, int& v) { v = std::atoi(x.get(t).c_str()); }
void read(XMLReader& x, const std::string& t, bool& v) { const std::string s = x.get(t); v = (s == "true" || s == "1"); }
void read(XMLReader& x, const std::string& t, Real& v) { v = Real(std::atof(x.get(t).c_str())); }
void read(XMLReader& x, const std::string& t, std::string& v) { v = x.get(t); }
static std::string wpath(XMLWriter& w, const std::string& t) {
  std::string p;
  for (size_t i = 0; i < w.stack.size(); ++i) p = XMLReader::join(p, w.stack[i]);
  return XMLReader::join(p, t);
}
void write(XMLWriter& w, const std::string& t, int v) { std::ostringstream s; s << v; w.kv[wpath(w, t)] = s.str(); }
void write(XMLWriter& w, const std::string& t, bool v) { w.kv[wpath(w, t)] = v ? "true" : "false"; }
void write(XMLWriter& w, const std::string& t, const Real& v) { std::ostringstream s; s.precision(17); s << toDouble(v); w.kv[wpath(w, t)] = s.str(); }
void write(XMLWriter& w, const std::string& t, const std::string& v) { w.kv[wpath(w, t)] = v; }
void push(XMLWriter& w, const std::string& t) { w.stack.push_back(t); }
void pop(XMLWriter& w) { w.stack.pop_back(); }
}  // namespace QDP

#include "io/aniso_io.h"
#include "actions/ferm/fermacts/clover_fermact_params_w.h"
namespace Chroma {
// makeFermCoeffs, lib/io/aniso_io.cc:63-80
multi1d<Real> makeFermCoeffs(const AnisoParam_t& aniso) {
  multi1d<Real> cf(Nd);
  cf = Real(1.0);
  if (aniso.anisoP)
    for (int mu = 0; mu < Nd; ++mu)
      if (mu != aniso.t_dir) cf[mu] = aniso.nu / aniso.xi_0;
  return cf;
}
// CloverFermActParams reader, clover_fermact_params_w.cc:27-99 (Mass | Kappa, clovCoeff | clovCoeffR + clovCoeffT, AnisoParam, TwistedM)
void read(XMLReader& xml, const std::string& path, CloverFermActParams& p) {
  XMLReader top(xml, path);
  p = CloverFermActParams();
  if (top.count("Mass")) read(top, "Mass", p.Mass);
  else if (top.count("Kappa")) { Real k; read(top, "Kappa", k); p.Mass = Real(1.0 / (2.0 * toDouble(k)) - 4.0); }
  else { QDPIO::cerr << "CloverFermActParams: neither Mass nor Kappa" << std::endl; QDP_abort(1); }
  if (top.count("AnisoParam")) {
    XMLReader a(top, "AnisoParam");
    read(a, "anisoP", p.anisoParam.anisoP); read(a, "t_dir", p.anisoParam.t_dir);
    read(a, "xi_0", p.anisoParam.xi_0); read(a, "nu", p.anisoParam.nu);
  }
  if (p.anisoParam.anisoP) { read(top, "clovCoeffR", p.clovCoeffR); read(top, "clovCoeffT", p.clovCoeffT); }   // :66-77
  else { read(top, "clovCoeff", p.clovCoeffR); p.clovCoeffT = p.clovCoeffR; }
  if (top.count("TwistedM")) { p.twisted_m_usedP = true; read(top, "TwistedM", p.twisted_m); }
}
void write(XMLWriter& xml, const std::string& path, const CloverFermActParams& p) {
  push(xml, path);
  write(xml, "Mass", p.Mass); write(xml, "clovCoeffR", p.clovCoeffR); write(xml, "clovCoeffT", p.clovCoeffT);
  if (p.twisted_m_usedP) write(xml, "TwistedM", p.twisted_m);
  pop(xml);
}
}  // namespace Chroma
