// A small FUNCTIONAL stand-in for the slice of the QDP++/Chroma API that (a) the B200 adapter (chroma_adapter/) and
// (b) the reference's own solver loops and clover site loops touch.  TEST INFRASTRUCTURE, never the product path.
// QDP++ is neither in /root/reference nor installed, so this restates its semantics for:
//   * scalar-site lattice types in cb2 site order: OLattice<T>::elem(site), Subset rb[2] / all,
//     subset assignment  x[s] = expr / += / -= / *= / = zero, norm2, innerProduct, gaussian, RNG::savern / setrn
//   * the word types Real / Double / RealF / RealD / ComplexF / ComplexD with the arithmetic the solver loops use
//   * RScalar / RComplex site arithmetic used by the clover site loops (clover_term_qdp_w.h:398-521, 619-815, 1562-1634)
// It lets oracle/Makefile compile lib/actions/ferm/invert/{invcg2,invbicgstab,minvcg2,reliable_cg}.cc UNMODIFIED from
// /root/reference (oracle/_ref/libref_chroma.so) and lets tests/adapter_exec.cc run the adapter's constructor and
// operator() against the real libb200clover.so.  Expressions are evaluated eagerly on whole fields; reductions run in
// site order in double precision, like scalar QDP++.
#ifndef MOCK_CHROMABASE_H
#define MOCK_CHROMABASE_H
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <map>
#include <string>
#include <vector>

#define START_CODE()
#define END_CODE()
#define QDP_ALIGN16

namespace QDP {
const int Nd = 4;
const int Nc = 3;
const int Ns = 4;

// ---------------------------------------------------------------------------------------------- site-level numbers
template <typename R> struct RScalar {
  R v;
  RScalar() : v(0) {}
  RScalar(R x) : v(x) {}
  R& elem() { return v; }
  const R& elem() const { return v; }
  RScalar& operator+=(const RScalar& o) { v += o.v; return *this; }
  RScalar& operator-=(const RScalar& o) { v -= o.v; return *this; }
};
template <typename R> inline RScalar<R> operator/(const RScalar<R>& a, const RScalar<R>& b) { return RScalar<R>(a.v / b.v); }
template <typename R> inline RScalar<R> operator*(const RScalar<R>& a, const RScalar<R>& b) { return RScalar<R>(a.v * b.v); }

template <typename R> struct RComplex {
  R re, im;
  RComplex() : re(0), im(0) {}
  RComplex(R a, R b) : re(a), im(b) {}
  R& real() { return re; }
  const R& real() const { return re; }
  R& imag() { return im; }
  const R& imag() const { return im; }
  RComplex& operator+=(const RComplex& o) { re += o.re; im += o.im; return *this; }
  RComplex& operator-=(const RComplex& o) { re -= o.re; im -= o.im; return *this; }
  RComplex& operator*=(const RScalar<R>& s) { re *= s.v; im *= s.v; return *this; }
  RComplex& operator/=(const RComplex& o) { *this = *this / o; return *this; }
};
template <typename R> inline RComplex<R> operator+(const RComplex<R>& a, const RComplex<R>& b) { return RComplex<R>(a.re + b.re, a.im + b.im); }
template <typename R> inline RComplex<R> operator-(const RComplex<R>& a, const RComplex<R>& b) { return RComplex<R>(a.re - b.re, a.im - b.im); }
template <typename R> inline RComplex<R> operator-(const RComplex<R>& a) { return RComplex<R>(-a.re, -a.im); }
template <typename R> inline RComplex<R> operator*(const RComplex<R>& a, const RComplex<R>& b) {
  return RComplex<R>(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
template <typename R> inline RComplex<R> operator*(const RComplex<R>& a, const RScalar<R>& s) { return RComplex<R>(a.re * s.v, a.im * s.v); }
template <typename R> inline RComplex<R> operator*(const RScalar<R>& s, const RComplex<R>& a) { return RComplex<R>(s.v * a.re, s.v * a.im); }
// complex division the way qdp_reality.h writes it: multiply by the reciprocal of |r|^2
template <typename R> inline RComplex<R> operator/(const RComplex<R>& l, const RComplex<R>& r) {
  const R tmp = R(1.0) / (r.re * r.re + r.im * r.im);
  return RComplex<R>((l.re * r.re + l.im * r.im) * tmp, (l.im * r.re - l.re * r.im) * tmp);
}
template <typename R> inline RScalar<R> real(const RComplex<R>& a) { return RScalar<R>(a.re); }
template <typename R> inline RScalar<R> imag(const RComplex<R>& a) { return RScalar<R>(a.im); }
template <typename R> inline RComplex<R> timesI(const RComplex<R>& a) { return RComplex<R>(-a.im, a.re); }
template <typename R> inline RComplex<R> adj(const RComplex<R>& a) { return RComplex<R>(a.re, -a.im); }
template <typename R> inline RComplex<R> conj(const RComplex<R>& a) { return RComplex<R>(a.re, -a.im); }
template <typename R> inline RComplex<R> cmplx(const RScalar<R>& a, const RScalar<R>& b) { return RComplex<R>(a.v, b.v); }
template <typename R> inline void zero_rep(RComplex<R>& a) { a.re = 0; a.im = 0; }

template <typename T> struct PScalar {
  T e;
  T& elem() { return e; }
  const T& elem() const { return e; }
};
template <typename T, int N> struct PColorVector {
  T c[N];
  T& elem(int i) { return c[i]; }
  const T& elem(int i) const { return c[i]; }
};
template <typename T, int N> struct PSpinVector {
  T s[N];
  T& elem(int i) { return s[i]; }
  const T& elem(int i) const { return s[i]; }
};
template <typename T, int N> struct PColorMatrix {
  T m[N * N];
  T& elem(int i, int j) { return m[N * i + j]; }
  const T& elem(int i, int j) const { return m[N * i + j]; }
};

// ---------------------------------------------------------------------------------------------- word-level scalars
// Real / Double / RealF / RealD (OScalar<PScalar<PScalar<RScalar<R>>>> in QDP++; .elem().elem().elem() works here too)
template <typename R> struct OReal {
  PScalar<PScalar<RScalar<R> > > e;
  OReal() {}
  OReal(double x) { e.e.e.v = (R)x; }
  template <typename R2> OReal(const OReal<R2>& o) { e.e.e.v = (R)o.val(); }
  R val() const { return e.e.e.v; }
  PScalar<PScalar<RScalar<R> > >& elem() { return e; }
  const PScalar<PScalar<RScalar<R> > >& elem() const { return e; }
  OReal& operator+=(const OReal& o) { e.e.e.v += o.val(); return *this; }
  OReal& operator/=(const OReal& o) { e.e.e.v /= o.val(); return *this; }
};
template <typename T> struct OScalar;     // only the spelling QDP++ code uses for Real: OScalar<PScalar<PScalar<RScalar<R>>>>
template <typename R> struct OScalar<PScalar<PScalar<RScalar<R> > > > : OReal<R> {
  OScalar() {}
  OScalar(double x) : OReal<R>(x) {}
  template <typename R2> OScalar(const OReal<R2>& o) : OReal<R>(o) {}
};
typedef OReal<float> RealF;
const double fuzz = 1.0e-5;              // QDP++'s "fuzz" constant (qdp_globalfuncs)
typedef OReal<double> RealD;
typedef RealD Real;      // a double-precision build of Chroma
typedef RealD Double;
typedef double REAL;
template <typename A, typename B> struct Promote { typedef double type; };
template <> struct Promote<float, float> { typedef float type; };
#define MOCK_REAL_BINOP(op)                                                                                          \
  template <typename A, typename B> inline OReal<typename Promote<A, B>::type> operator op(const OReal<A>& a, const OReal<B>& b) { \
    typedef typename Promote<A, B>::type P; return OReal<P>((P)a.val() op (P)b.val()); }                             \
  template <typename A> inline OReal<A> operator op(const OReal<A>& a, double b) { return OReal<A>(a.val() op (A)b); } \
  template <typename A> inline OReal<A> operator op(double a, const OReal<A>& b) { return OReal<A>((A)a op b.val()); }
MOCK_REAL_BINOP(+)
MOCK_REAL_BINOP(-)
MOCK_REAL_BINOP(*)
MOCK_REAL_BINOP(/)
#undef MOCK_REAL_BINOP
template <typename A> inline OReal<A> operator-(const OReal<A>& a) { return OReal<A>(-a.val()); }
struct Boolean { bool b; };
#define MOCK_REAL_CMP(op)                                                                                              \
  template <typename A, typename B> inline Boolean operator op(const OReal<A>& a, const OReal<B>& b) { Boolean r = {(double)a.val() op (double)b.val()}; return r; } \
  template <typename A> inline Boolean operator op(const OReal<A>& a, double b) { Boolean r = {(double)a.val() op b}; return r; }
MOCK_REAL_CMP(<)
MOCK_REAL_CMP(>)
MOCK_REAL_CMP(<=)
MOCK_REAL_CMP(>=)
MOCK_REAL_CMP(==)
#undef MOCK_REAL_CMP
inline Boolean operator&&(const Boolean& a, const Boolean& b) { Boolean r = {a.b && b.b}; return r; }
inline Boolean operator||(const Boolean& a, const Boolean& b) { Boolean r = {a.b || b.b}; return r; }
inline Boolean operator!(const Boolean& a) { Boolean r = {!a.b}; return r; }
inline bool toBool(const Boolean& b) { return b.b; }
inline bool toBool(bool b) { return b; }
template <typename A> inline double toDouble(const OReal<A>& r) { return (double)r.val(); }
template <typename A> inline float toFloat(const OReal<A>& r) { return (float)r.val(); }
template <typename A> inline OReal<A> sqrt(const OReal<A>& r) { return OReal<A>(std::sqrt(r.val())); }
template <typename A> inline OReal<A> fabs(const OReal<A>& r) { return OReal<A>(std::fabs(r.val())); }
template <typename A> inline std::ostream& operator<<(std::ostream& o, const OReal<A>& r) { return o << r.val(); }

template <typename R> struct OComplex {
  R re, im;
  OComplex() : re(0), im(0) {}
  OComplex(R a, R b) : re(a), im(b) {}
  template <typename R2> OComplex(const OComplex<R2>& o) : re((R)o.re), im((R)o.im) {}
  template <typename R2> OComplex(const OReal<R2>& o) : re((R)o.val()), im(0) {}
  template <typename R2> OComplex& operator/=(const OReal<R2>& d) { re = (R)(re / d.val()); im = (R)(im / d.val()); return *this; }
};
typedef OComplex<float> ComplexF;
typedef OComplex<double> ComplexD;
typedef ComplexD Complex;
typedef ComplexD DComplex;
template <typename R> inline OReal<R> real(const OComplex<R>& c) { return OReal<R>(c.re); }
template <typename R> inline OReal<R> imag(const OComplex<R>& c) { return OReal<R>(c.im); }
template <typename R> inline OComplex<R> operator*(const OComplex<R>& a, const OComplex<R>& b) { return OComplex<R>(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
template <typename R> inline OComplex<R> operator+(const OComplex<R>& a, const OComplex<R>& b) { return OComplex<R>(a.re + b.re, a.im + b.im); }
template <typename R> inline OComplex<R> operator-(const OComplex<R>& a, const OComplex<R>& b) { return OComplex<R>(a.re - b.re, a.im - b.im); }
template <typename R> inline OComplex<R> operator/(const OComplex<R>& l, const OComplex<R>& r) {
  const R tmp = R(1.0) / (r.re * r.re + r.im * r.im);
  return OComplex<R>((l.re * r.re + l.im * r.im) * tmp, (l.im * r.re - l.re * r.im) * tmp);
}
template <typename R> inline OComplex<R> conj(const OComplex<R>& a) { return OComplex<R>(a.re, -a.im); }
template <typename R> inline OReal<R> norm2(const OComplex<R>& a) { return OReal<R>(a.re * a.re + a.im * a.im); }
template <typename R> inline std::ostream& operator<<(std::ostream& o, const OComplex<R>& c) { return o << "(" << c.re << "," << c.im << ")"; }

// ---------------------------------------------------------------------------------------------- containers, layout
template <typename T> class multi1d {
 public:
  multi1d() {}
  explicit multi1d(int n) : d(n) {}
  int size() const { return (int)d.size(); }
  void resize(int n) { d.resize(n); }
  T& operator[](int i) { return d[i]; }
  const T& operator[](int i) const { return d[i]; }
  const T* slice() const { return d.data(); }
  multi1d& operator=(const T& x) { for (size_t i = 0; i < d.size(); ++i) d[i] = x; return *this; }
 private:
  std::vector<T> d;
};
template <> class multi1d<bool> {     // std::vector<bool> has no addressable elements
 public:
  multi1d() {}
  explicit multi1d(int n) : d(n, 0) {}
  int size() const { return (int)d.size(); }
  void resize(int n) { d.resize(n); }
  bool& operator[](int i) { return reinterpret_cast<bool&>(d[i]); }
  const bool& operator[](int i) const { return reinterpret_cast<const bool&>(d[i]); }
 private:
  std::vector<unsigned char> d;
};

template <typename T> class multi2d {     // multi2d<T> z(n2, n1); z[j][i]
 public:
  multi2d() : n1(0) {}
  multi2d(int n2_, int n1_) : n1(n1_), d((size_t)n2_ * n1_) {}
  T* operator[](int j) { return d.data() + (size_t)j * n1; }
  const T* operator[](int j) const { return d.data() + (size_t)j * n1; }
 private:
  int n1;
  std::vector<T> d;
};

namespace Layout {
// test harness: the LOCAL lattice of this rank, the process grid and this rank's coordinate in it
void mockSetup(const int local_dims[4], const int grid[4], const int coord[4], int node, int nodes);
const multi1d<int>& lattSize();      // GLOBAL extents
const multi1d<int>& subgridLattSize();
const multi1d<int>& logicalSize();
const multi1d<int>& nodeCoord();
int nodeNumber();
int numNodes();
int sitesOnNode();
int vol();
}

class Subset {
 public:
  Subset() : lo(0), hi(0) {}
  void make(int lo_, int hi_) { lo = lo_; hi = hi_; tab.resize(hi - lo); for (int i = lo; i < hi; ++i) tab[i - lo] = i; }
  int start() const { return lo; }
  int end() const { return hi - 1; }
  int numSiteTable() const { return hi - lo; }
  const multi1d<int>& siteTable() const { return tab; }
  bool hasOrderedRep() const { return true; }
 private:
  int lo, hi;
  multi1d<int> tab;
};
extern Subset all;
extern Subset rb[2];
struct Zero {};
extern Zero zero;

// ---------------------------------------------------------------------------------------------- lattice fields
template <typename L> struct SubsetProxy;
template <typename T> struct SiteWord { typedef typename T::Word type; };
template <typename R> struct SiteWord<PScalar<PScalar<RScalar<R> > > > { typedef R type; };   // LatticeReal
template <typename T> class OLattice {
 public:
  typedef T Site;
  OLattice() : d(Layout::sitesOnNode()) {}
  OLattice(const Zero&) : d(Layout::sitesOnNode()) {}
  OLattice& operator=(const Zero&) { std::fill(d.begin(), d.end(), T()); return *this; }
  T& elem(int i) { return d[i]; }
  const T& elem(int i) const { return d[i]; }
  SubsetProxy<OLattice> operator[](const Subset& s) { SubsetProxy<OLattice> p = {*this, s}; return p; }
  // flat view in words of the site type
  typedef typename SiteWord<T>::type Word;
  Word* words() { return reinterpret_cast<Word*>(d.data()); }
  const Word* words() const { return reinterpret_cast<const Word*>(d.data()); }
  static int wordsPerSite() { return (int)(sizeof(T) / sizeof(Word)); }
 private:
  std::vector<T> d;
};
template <typename R, int NSPIN> struct FermSite : PSpinVector<PColorVector<RComplex<R>, Nc>, NSPIN> {
  typedef R Word;
  FermSite() { for (int s = 0; s < NSPIN; ++s) for (int c = 0; c < Nc; ++c) this->s[s].c[c] = RComplex<R>(); }
};
template <typename R> struct CMSite : PScalar<PColorMatrix<RComplex<R>, Nc> > {
  typedef R Word;
  CMSite() { for (int i = 0; i < Nc * Nc; ++i) this->e.m[i] = RComplex<R>(); }
};
template <typename R> struct RealSite : PScalar<PScalar<RScalar<R> > > { typedef R Word; };
typedef OLattice<FermSite<float, Ns> > LatticeFermionF;
typedef OLattice<FermSite<double, Ns> > LatticeFermionD;
typedef LatticeFermionD LatticeFermion;
typedef OLattice<FermSite<float, 1> > LatticeStaggeredFermionF;
typedef OLattice<FermSite<double, 1> > LatticeStaggeredFermionD;
typedef LatticeStaggeredFermionD LatticeStaggeredFermion;
typedef OLattice<CMSite<float> > LatticeColorMatrixF;
typedef OLattice<CMSite<double> > LatticeColorMatrixD;
typedef LatticeColorMatrixD LatticeColorMatrix;
template <typename T> struct WordType;
template <typename S> struct WordType<OLattice<S> > { typedef typename SiteWord<S>::type Type_t; };

// eager "expressions": every operator returns a whole field
#define MOCK_FOR_WORDS(L, x) const int n_ = Layout::sitesOnNode() * L::wordsPerSite(); for (int x = 0; x < n_; ++x)
template <typename S> inline OLattice<S> operator+(const OLattice<S>& a, const OLattice<S>& b) {
  OLattice<S> r; MOCK_FOR_WORDS(OLattice<S>, i) r.words()[i] = a.words()[i] + b.words()[i]; return r;
}
template <typename S> inline OLattice<S> operator-(const OLattice<S>& a, const OLattice<S>& b) {
  OLattice<S> r; MOCK_FOR_WORDS(OLattice<S>, i) r.words()[i] = a.words()[i] - b.words()[i]; return r;
}
template <typename S, typename R> inline OLattice<S> operator*(const OReal<R>& a, const OLattice<S>& b) {
  typedef typename SiteWord<S>::type W; const W s = (W)a.val();
  OLattice<S> r; MOCK_FOR_WORDS(OLattice<S>, i) r.words()[i] = s * b.words()[i]; return r;
}
template <typename S, typename R> inline OLattice<S> operator*(const OLattice<S>& b, const OReal<R>& a) { return a * b; }
template <typename S, typename R> inline OLattice<S> operator*(const OComplex<R>& a, const OLattice<S>& b) {
  typedef typename SiteWord<S>::type W; const W ar = (W)a.re, ai = (W)a.im;
  OLattice<S> r;
  const int n = Layout::sitesOnNode() * OLattice<S>::wordsPerSite() / 2;
  for (int i = 0; i < n; ++i) {
    const W br = b.words()[2 * i], bi = b.words()[2 * i + 1];
    r.words()[2 * i] = ar * br - ai * bi; r.words()[2 * i + 1] = ar * bi + ai * br;
  }
  return r;
}
template <typename L> struct SubsetProxy {
  L& l; const Subset& s;
  typedef typename L::Word W;
  int w0() const { return s.start() * L::wordsPerSite(); }
  int w1() const { return (s.end() + 1) * L::wordsPerSite(); }
  SubsetProxy& operator=(const L& o) { for (int i = w0(); i < w1(); ++i) l.words()[i] = o.words()[i]; return *this; }
  // precision-converting assignment (LatticeFermionF <-> LatticeFermionD, reliable_cg.cc:58,124)
  template <typename S2> SubsetProxy& operator=(const OLattice<S2>& o) { for (int i = w0(); i < w1(); ++i) l.words()[i] = (W)o.words()[i]; return *this; }
  template <typename S2> SubsetProxy& operator+=(const OLattice<S2>& o) { for (int i = w0(); i < w1(); ++i) l.words()[i] += (W)o.words()[i]; return *this; }
  SubsetProxy& operator+=(const L& o) { for (int i = w0(); i < w1(); ++i) l.words()[i] += o.words()[i]; return *this; }
  SubsetProxy& operator-=(const L& o) { for (int i = w0(); i < w1(); ++i) l.words()[i] -= o.words()[i]; return *this; }
  template <typename R> SubsetProxy& operator*=(const OReal<R>& a) { const W x = (W)a.val(); for (int i = w0(); i < w1(); ++i) l.words()[i] *= x; return *this; }
  SubsetProxy& operator=(const Zero&) { for (int i = w0(); i < w1(); ++i) l.words()[i] = 0; return *this; }
};
// colour-matrix fields: products, adjoint, nearest-neighbour shifts (what mesField needs, lib/meas/glue/mesfield.cc:30-78)
enum { BACKWARD = -1, FORWARD = 1 };
namespace Layout { const int* mockNeighbourTable(int dir, int mu); }    // [sitesOnNode]: site of x + dir*mu, periodic, cb2 order
template <typename R> inline OLattice<CMSite<R> > operator*(const OLattice<CMSite<R> >& a, const OLattice<CMSite<R> >& b) {
  OLattice<CMSite<R> > r;
  const int n = Layout::sitesOnNode();
  for (int s = 0; s < n; ++s)
    for (int i = 0; i < Nc; ++i)
      for (int j = 0; j < Nc; ++j) {
        RComplex<R> acc = a.elem(s).elem().elem(i, 0) * b.elem(s).elem().elem(0, j);
        for (int k = 1; k < Nc; ++k) acc += a.elem(s).elem().elem(i, k) * b.elem(s).elem().elem(k, j);
        r.elem(s).elem().elem(i, j) = acc;
      }
  return r;
}
template <typename R> inline OLattice<CMSite<R> > adj(const OLattice<CMSite<R> >& a) {
  OLattice<CMSite<R> > r;
  const int n = Layout::sitesOnNode();
  for (int s = 0; s < n; ++s)
    for (int i = 0; i < Nc; ++i)
      for (int j = 0; j < Nc; ++j) r.elem(s).elem().elem(i, j) = adj(a.elem(s).elem().elem(j, i));
  return r;
}
template <typename S> inline OLattice<S> shift(const OLattice<S>& a, int dir, int mu) {
  OLattice<S> r;
  const int* nb = Layout::mockNeighbourTable(dir, mu);
  const int n = Layout::sitesOnNode();
  for (int s = 0; s < n; ++s) r.elem(s) = a.elem(nb[s]);
  return r;
}
template <typename S> inline OLattice<S>& operator+=(OLattice<S>& a, const OLattice<S>& b) { MOCK_FOR_WORDS(OLattice<S>, i) a.words()[i] += b.words()[i]; return a; }
template <typename S> inline OLattice<S>& operator-=(OLattice<S>& a, const OLattice<S>& b) { MOCK_FOR_WORDS(OLattice<S>, i) a.words()[i] -= b.words()[i]; return a; }
template <typename S, typename R> inline OLattice<S>& operator*=(OLattice<S>& a, const OReal<R>& f) {
  typedef typename SiteWord<S>::type W; const W x = (W)f.val(); MOCK_FOR_WORDS(OLattice<S>, i) a.words()[i] *= x; return a;
}

// reductions: site order, double accumulation (scalar QDP++)
template <typename S> inline Double norm2(const OLattice<S>& a, const Subset& s) {
  double acc = 0;
  const int w = OLattice<S>::wordsPerSite();
  for (int i = s.start() * w; i < (s.end() + 1) * w; ++i) acc += (double)a.words()[i] * (double)a.words()[i];
  return Double(acc);
}
template <typename S> inline Double norm2(const OLattice<S>& a) { return norm2(a, all); }
template <typename S> inline ComplexD innerProduct(const OLattice<S>& a, const OLattice<S>& b, const Subset& s) {
  double re = 0, im = 0;
  const int w = OLattice<S>::wordsPerSite();
  for (int i = s.start() * w / 2; i < (s.end() + 1) * w / 2; ++i) {
    const double ar = a.words()[2 * i], ai = a.words()[2 * i + 1], br = b.words()[2 * i], bi = b.words()[2 * i + 1];
    re += ar * br + ai * bi; im += ar * bi - ai * br;       // conj(a) * b
  }
  return ComplexD(re, im);
}
template <typename S> inline Double innerProductReal(const OLattice<S>& a, const OLattice<S>& b, const Subset& s) { return real(innerProduct(a, b, s)); }

// the RNG: a 64-bit LCG whose state can be saved and restored like QDP::RNG's
typedef uint64_t Seed;
namespace RNG {
void savern(Seed& s);
void setrn(const Seed& s);
double gauss();
}
template <typename S> inline void gaussian(OLattice<S>& x, const Subset& s) {
  const int w = OLattice<S>::wordsPerSite();
  for (int i = s.start() * w; i < (s.end() + 1) * w; ++i) x.words()[i] = (typename SiteWord<S>::type)RNG::gauss();
}

namespace Hints {
template <typename T> inline void moveToFastMemoryHint(T&, bool = false) {}
template <typename T> inline void revertFromFastMemoryHint(T&, bool = false) {}
}
using namespace Hints;

namespace QDPInternal {
// sums over the ranks of the test harness (tests/mock_chroma/mock_qdp.cc: a shared-memory segment between forked ranks)
void globalSumArray(int* a, int n);
void globalSum(int& x);
}
namespace QDPIO { extern std::ostream& cout; extern std::ostream& cerr; }
void QDP_abort(int);
void QDP_error_exit(const char* fmt, ...);

// A flat key -> value store stands in for libxml2: XMLReader(top, "path") narrows the prefix.
class XMLReader {
 public:
  XMLReader() {}
  explicit XMLReader(const std::map<std::string, std::string>& kv_) : kv(kv_) {}
  XMLReader(XMLReader& top, const std::string& path) : kv(top.kv), prefix(join(top.prefix, path)) {}
  int count(const std::string& tag) const {
    const std::string k = join(prefix, tag);
    for (std::map<std::string, std::string>::const_iterator it = kv.begin(); it != kv.end(); ++it)
      if (it->first == k || it->first.compare(0, k.size() + 1, k + "/") == 0) return 1;
    return 0;
  }
  std::string get(const std::string& tag) const;
  static std::string join(const std::string& a, const std::string& b) {
    std::string t = b; while (!t.empty() && (t[0] == '/' || t[0] == '.')) t.erase(0, 1);
    return a.empty() ? t : (t.empty() ? a : a + "/" + t);
  }
  std::map<std::string, std::string> kv;
  std::string prefix;
};
class XMLWriter { public: std::map<std::string, std::string> kv; std::vector<std::string> stack; };
void read(XMLReader&, const std::string&, int&);
void read(XMLReader&, const std::string&, bool&);
void read(XMLReader&, const std::string&, Real&);
void read(XMLReader&, const std::string&, std::string&);
void write(XMLWriter&, const std::string&, int);
void write(XMLWriter&, const std::string&, bool);
void write(XMLWriter&, const std::string&, const Real&);
void write(XMLWriter&, const std::string&, const std::string&);
void push(XMLWriter&, const std::string&);
void pop(XMLWriter&);

class StopWatch {
 public:
  void reset() {} void start() {} void stop() {}
  double getTimeInSeconds() const { return 0.0; }
};
class FlopCounter {
 public:
  void reset() {}
  void addSiteFlops(unsigned long, const Subset&) {}
  void addSiteFlops(unsigned long) {}
  void addFlops(unsigned long) {}
  void report(const std::string&, double) {}
};
}  // namespace QDP

namespace Chroma {
using namespace QDP;
enum PlusMinus { PLUS = 1, MINUS = -1 };
struct StringFactoryError {};
}
#endif
