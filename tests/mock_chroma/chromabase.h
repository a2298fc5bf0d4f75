// Minimal stand-in for the slice of the QDP++/Chroma API the B200 adapter touches.
// TEST INFRASTRUCTURE: lets tests/test_adapter_compiles.py type-check chroma_adapter/*.cc without QDP++ (which is
// neither in /root/reference nor installed).  Signatures follow the uses in quda_solvers/syssolver_linop_clover_quda_w.h
// and lwldslash_w_cppd.cc; nothing here computes anything.
#ifndef MOCK_CHROMABASE_H
#define MOCK_CHROMABASE_H
#include <cmath>
#include <iostream>
#include <string>
#include <vector>

#define START_CODE()
#define END_CODE()

namespace QDP {
const int Nd = 4;
typedef double REAL;
struct Real { double v; Real(double x = 0) : v(x) {} };
typedef Real Double;
inline Real operator*(const Real& a, const Real& b) { return Real(a.v * b.v); }
inline Real operator/(const Real& a, const Real& b) { return Real(a.v / b.v); }
struct Boolean { bool b; };
inline Boolean operator>(const Real& a, const Real& b) { Boolean r = {a.v > b.v}; return r; }
inline Real operator+(const Real& a, const Real& b) { return Real(a.v + b.v); }
inline double toDouble(const Real& r) { return r.v; }
inline bool toBool(const Boolean& b) { return b.b; }
inline Real sqrt(const Real& r) { return Real(std::sqrt(r.v)); }
inline std::ostream& operator<<(std::ostream& o, const Real& r) { return o << r.v; }

template <typename T> class multi1d {
 public:
  multi1d() {}
  explicit multi1d(int n) : d(n) {}
  int size() const { return (int)d.size(); }
  void resize(int n) { d.resize(n); }
  T& operator[](int i) { return d[i]; }
  const T& operator[](int i) const { return d[i]; }
 private:
  std::vector<T> d;
};

struct Subset { int start() const { return 0; } };
extern Subset all;
extern Subset rb[2];
struct Zero {};
extern Zero zero;

struct RComplexRef { REAL re, im; REAL& real() { return re; } const REAL& real() const { return re; } };
struct ColorVec { RComplexRef c[3]; RComplexRef& elem(int i) { return c[i]; } const RComplexRef& elem(int i) const { return c[i]; } };
struct SpinVec { ColorVec s[4]; ColorVec& elem(int i) { return s[i]; } const ColorVec& elem(int i) const { return s[i]; } };
struct ColorMat { RComplexRef m[9]; RComplexRef& elem(int i, int j) { return m[3 * i + j]; } const RComplexRef& elem(int i, int j) const { return m[3 * i + j]; } };
struct ScalarCM { ColorMat m; ColorMat& elem() { return m; } const ColorMat& elem() const { return m; } };

template <typename L> struct SubsetProxy {
  L& l;
  SubsetProxy& operator=(const L&) { return *this; }
  SubsetProxy& operator-=(const L&) { return *this; }
  SubsetProxy& operator*=(const Real&) { return *this; }   // chi[rb[0]] *= mhalf, seoprec_clover_linop_w.cc:113
  SubsetProxy& operator=(const Zero&) { return *this; }
};
class LatticeFermion {
 public:
  LatticeFermion() {}
  LatticeFermion(const Zero&) {}
  LatticeFermion& operator=(const Zero&) { return *this; }
  SpinVec& elem(int) { return site; }
  const SpinVec& elem(int) const { return site; }
  SubsetProxy<LatticeFermion> operator[](const Subset&) { SubsetProxy<LatticeFermion> p = {*this}; return p; }
 private:
  SpinVec site;
};
class LatticeColorMatrix {
 public:
  ScalarCM& elem(int) { return site; }
  const ScalarCM& elem(int) const { return site; }
 private:
  ScalarCM site;
};
template <typename T> struct WordType { typedef REAL Type_t; };
inline Double norm2(const LatticeFermion&, const Subset&) { return Double(1.0); }
inline void gaussian(LatticeFermion&, const Subset&) {}

namespace Layout {
const multi1d<int>& lattSize();
const multi1d<int>& logicalSize();
const multi1d<int>& nodeCoord();
int nodeNumber();
int numNodes();
}
namespace QDPInternal {
void globalSumArray(int* a, int n);
void globalSum(int& x);
}
namespace QDPIO { extern std::ostream& cout; extern std::ostream& cerr; }
void QDP_abort(int);

class XMLReader {
 public:
  XMLReader() {}
  XMLReader(XMLReader&, const std::string&) {}
  int count(const std::string&) const { return 0; }
};
class XMLWriter {};
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
}  // namespace QDP

namespace Chroma {
using namespace QDP;
enum PlusMinus { PLUS = 1, MINUS = -1 };
struct StringFactoryError {};
}
#endif
