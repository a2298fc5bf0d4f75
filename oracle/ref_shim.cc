// ref_shim.cc -- extern "C" door onto the reference's OWN scalar Dslash / CloverSchur4D.
//
// TEST INFRASTRUCTURE ONLY (see oracle/oracle.c header).  This file is ours; the objects it
// links against are compiled, unmodified and in place, from
//   /root/reference/other_libs/cpp_wilson_dslash/lib/{shift_table_scalar,cpp_dslash_scalar_64bit,
//   cpp_dslash_scalar_32bit,cpp_clover_scalar_64bit,cpp_clover_scalar_32bit,dispatch_scalar_openmp}.cc
// by oracle/Makefile into oracle/_ref/libref_dslash.so.  No reference source is copied.
//
// The reference takes its geometry as three C callbacks (include/cpp_dslash_scalar.h:39-43);
// we hand it QDP++'s cb2 layout, which is also what its own loops assume
// (lib/shift_table_scalar.cc:155-214).
#include <cpp_dslash_scalar.h>
#include <cpp_clover_scalar.h>
#include <cstddef>

using namespace CPlusPlusWilsonDslash;
using namespace CPlusPlusClover;

namespace {
int gL[4] = {0, 0, 0, 0};

int linearSiteIndex(const int c[]) {
  int V = gL[0] * gL[1] * gL[2] * gL[3];
  int cb = (c[0] + c[1] + c[2] + c[3]) & 1;
  return ((c[3] * gL[2] + c[2]) * gL[1] + c[1]) * (gL[0] / 2) + c[0] / 2 + cb * (V / 2);
}
void siteCoords(int c[], int /*node*/, int idx) {
  int V = gL[0] * gL[1] * gL[2] * gL[3], Vh = V / 2, Lxh = gL[0] / 2;
  int cb = idx / Vh, r = idx % Vh;
  int xh = r % Lxh; r /= Lxh;
  c[1] = r % gL[1]; r /= gL[1];
  c[2] = r % gL[2]; r /= gL[2];
  c[3] = r;
  c[0] = 2 * xh + ((cb + c[1] + c[2] + c[3]) & 1);
}
int nodeNum(const int[]) { return 0; }
void setL(const int L[4]) { for (int i = 0; i < 4; ++i) gL[i] = L[i]; }
}  // namespace

extern "C" {

void* ref_dslash_create_d(const int L[4]) { setL(L); return new Dslash<double>(L, siteCoords, linearSiteIndex, nodeNum); }
void ref_dslash_free_d(void* h) { delete static_cast<Dslash<double>*>(h); }
// cb is the SOURCE checkerboard, exactly as Dslash<double>::operator() takes it.
void ref_dslash_apply_d(void* h, double* res, double* psi, double* packed_u, int isign, int cb) {
  (*static_cast<Dslash<double>*>(h))(res, psi, packed_u, isign, cb);
}

void* ref_dslash_create_f(const int L[4]) { setL(L); return new Dslash<float>(L, siteCoords, linearSiteIndex, nodeNum); }
void ref_dslash_free_f(void* h) { delete static_cast<Dslash<float>*>(h); }
void ref_dslash_apply_f(void* h, float* res, float* psi, float* packed_u, int isign, int cb) {
  (*static_cast<Dslash<float>*>(h))(res, psi, packed_u, isign, cb);
}

// CloverSchur4D: res_o = clov_oo psi_o - D_oe invclov_ee D_eo psi_o (no 1/4: the caller folds 1/2 into the
// links).  clov arrays are CloverTerm[V] = 2 x {diag[8], off_diag[16][2]} = 80 reals per site.
// NOTE: its two site loops run inside one OpenMP region with no barrier between them
// (lib/dispatch_scalar_openmp.cc:63-72, lib/cpp_clover_scalar_64bit.cc:137-248), so it is only
// deterministic with OMP_NUM_THREADS=1; callers of this entry must set that.
void* ref_clover_create_d(const int L[4]) { setL(L); return new CloverSchur4D<double>(L, siteCoords, linearSiteIndex, nodeNum); }
void ref_clover_free_d(void* h) { delete static_cast<CloverSchur4D<double>*>(h); }
void ref_clover_apply_d(void* h, double* res, const double* psi, const double* packed_u,
                        const double* clov_oo, const double* invclov_ee, int isign) {
  (*static_cast<CloverSchur4D<double>*>(h))(res, psi, packed_u, clov_oo, invclov_ee, isign);
}

}  // extern "C"
