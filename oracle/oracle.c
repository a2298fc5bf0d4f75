/*
 * oracle.c -- CPU restatement of Chroma's even-odd preconditioned Wilson-clover path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under chroma_b200/ may call, link or load this
 * file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
 * arm use it, and only as the checker / CPU baseline -- never as the product path.
 *
 * Every function cites the reference lines (relative to /root/reference) it follows.
 * Data layouts are the ones QDP++ hands to the solver plugin:
 *   site order   "cb2": idx = cb*Vh + ((t*Lz+z)*Ly+y)*(Lx/2) + x/2, cb=(x+y+z+t)&1
 *                (other_libs/cpp_wilson_dslash/lib/shift_table_scalar.cc:155-214)
 *   fermion      double[V][spin 4][colour 3][re,im]
 *   gauge        four arrays u[mu] = double[V][row 3][col 3][re,im]
 *   packed gauge double[V][mu 4][col 3][row 3][re,im]  (transposed links,
 *                other_libs/cpp_wilson_dslash/lib/qdp_packer_nopad.cc:13-19)
 *   clover       per site: diag[2][6] reals then offd[2][15] complex = 72 reals
 *                (lib/actions/ferm/linop/clover_term_qdp_w.h:19-24)
 *
 * Parity pinning (tests/test_oracle.py, tests/test_golden.py): every restatement in this file is compared with the
 * reference's OWN code compiled unmodified from /root/reference (oracle/Makefile -> oracle/_ref/):
 *   orc_dslash, the composed Schur operator   Dslash<double|float>, CloverSchur4D<double>       (_ref/libref_dslash.so)
 *   orc_mesfield, orc_make_clov               mesField, makeClovSiteLoop: bit for bit            (_ref/libref_chroma.so)
 *   orc_ldagdlinv, orc_clover_apply           LDagDLInvSiteLoop (1 ulp), applySiteLoop (bit for bit)
 *   orc_invcg2, orc_invbicgstab, orc_minvcg2  InvCG2, InvBiCGStab, MInvCG2: same counts, solutions to 1e-15
 *   reliable CG / BiCGStab (oracle.py)        InvCGReliable, InvBiCGStabReliable with an fp32 inner operator
 * The Chroma-level sources compile against tests/mock_chroma, a functional stand-in for the slice of QDP++ they touch
 * (QDP++ itself is not available).  The symmetric operator, the twisted-mass term and the qprop decompositions are
 * compositions of those pinned pieces (their own source files instantiate all of Chroma) and are checked through
 * properties: A_oo^-1 M_asym, <chi,M psi> = <M^dag chi,psi>, explicit Gamma(15) algebra, the unpreconditioned system.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NC 3
#define SPINOR 24   /* reals per site */
#define LINK 18
#define CLOV 72

typedef struct { double re, im; } cplx;

static inline cplx c_mul(cplx a, cplx b) { cplx r = { a.re*b.re - a.im*b.im, a.re*b.im + a.im*b.re }; return r; }
static inline cplx c_conj(cplx a) { cplx r = { a.re, -a.im }; return r; }
static inline cplx c_add(cplx a, cplx b) { cplx r = { a.re+b.re, a.im+b.im }; return r; }
static inline cplx c_sub(cplx a, cplx b) { cplx r = { a.re-b.re, a.im-b.im }; return r; }
static inline cplx c_timesI(cplx a) { cplx r = { -a.im, a.re }; return r; }
static inline cplx c_neg(cplx a) { cplx r = { -a.re, -a.im }; return r; }
/* complex division as std::complex / RComplex do it (textbook formula) */
static inline cplx c_div(cplx a, cplx b) {
  double d = b.re*b.re + b.im*b.im;
  cplx r = { (a.re*b.re + a.im*b.im)/d, (a.im*b.re - a.re*b.im)/d };
  return r;
}

/* ------------------------------------------------------------------------- */
/* Geometry: QDP++ cb2 layout, restated from shift_table_scalar.cc:155-214    */
/* ------------------------------------------------------------------------- */
int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* bench.py --impl reference runs under torchrun, which exports OMP_NUM_THREADS=1: the CPU arm sets its team itself
 * (the reference's OpenMP dispatcher in oracle/_ref shares this libgomp, so it follows). */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int orc_site_index(const int L[4], const int c[4]) {
  int V = L[0]*L[1]*L[2]*L[3];
  int cb = (c[0]+c[1]+c[2]+c[3]) & 1;
  return ((c[3]*L[2]+c[2])*L[1]+c[1])*(L[0]/2) + c[0]/2 + cb*(V/2);
}

void orc_site_coords(const int L[4], int idx, int c[4]) {
  int V = L[0]*L[1]*L[2]*L[3], Vh = V/2, Lxh = L[0]/2;
  int cb = idx / Vh, r = idx % Vh;
  int xh = r % Lxh; r /= Lxh;
  c[1] = r % L[1]; r /= L[1];
  c[2] = r % L[2]; r /= L[2];
  c[3] = r;
  c[0] = 2*xh + ((cb + c[1] + c[2] + c[3]) & 1);
}

static inline int nbr(const int L[4], const int c[4], int mu, int dir) {
  int n[4] = { c[0], c[1], c[2], c[3] };
  n[mu] = (c[mu] + dir + L[mu]) % L[mu];
  return orc_site_index(L, n);
}

/* neighbour tables: fwd[4*site+mu], bwd[4*site+mu] (shift_table_scalar.cc:118-147) */
static void build_tables(const int L[4], int** fwd_out, int** bwd_out) {
  int V = L[0]*L[1]*L[2]*L[3];
  int* fwd = (int*)malloc(sizeof(int)*4*(size_t)V);
  int* bwd = (int*)malloc(sizeof(int)*4*(size_t)V);
#pragma omp parallel for
  for (int s = 0; s < V; ++s) {
    int c[4]; orc_site_coords(L, s, c);
    for (int mu = 0; mu < 4; ++mu) { fwd[4*s+mu] = nbr(L, c, mu, +1); bwd[4*s+mu] = nbr(L, c, mu, -1); }
  }
  *fwd_out = fwd; *bwd_out = bwd;
}

/* ------------------------------------------------------------------------- */
/* Boundary phases and anisotropy factors folded into the links               */
/* ------------------------------------------------------------------------- */
/* lib/actions/ferm/fermbcs/simple_fermbc.h:87-103: u[m] *= boundary[m] on x_m = L_m-1 */
void orc_apply_fermbc(const int L[4], double* const u[4], const int boundary[4]) {
  int V = L[0]*L[1]*L[2]*L[3];
  for (int mu = 0; mu < 4; ++mu) {
    if (boundary[mu] == 1) continue;
#pragma omp parallel for
    for (int s = 0; s < V; ++s) {
      int c[4]; orc_site_coords(L, s, c);
      if (c[mu] == L[mu]-1)
        for (int k = 0; k < LINK; ++k) u[mu][(size_t)s*LINK+k] *= (double)boundary[mu];
    }
  }
}

/* lwldslash_w_cppd.cc:115-121 (u[mu] *= coeffs[mu]) followed by
 * qdp_packer_nopad.cc:13-19 (u_tmp[mu+4*ix] = transpose(u[mu](ix))) */
void orc_pack_gauge(const int L[4], const double* const u[4], const double coeffs[4], double* packed) {
  int V = L[0]*L[1]*L[2]*L[3];
#pragma omp parallel for
  for (int s = 0; s < V; ++s)
    for (int mu = 0; mu < 4; ++mu) {
      const double* in = u[mu] + (size_t)s*LINK;
      double* out = packed + ((size_t)s*4 + mu)*LINK;
      for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) {
        out[(c*3+r)*2+0] = coeffs[mu]*in[(r*3+c)*2+0];
        out[(c*3+r)*2+1] = coeffs[mu]*in[(r*3+c)*2+1];
      }
    }
}

/* ------------------------------------------------------------------------- */
/* Wilson hopping term                                                        */
/* ------------------------------------------------------------------------- */
/* su3_mult: res[s][row] = sum_col u[col][row]*h[s][col]  (cpp_dslash_matvec64bit_c.h:14-150)
 * su3_adj_mult: res[s][row] = sum_col conj(u[row][col])*h[s][col]  (:152-296)
 * with u the PACKED (transposed) link. */
static inline void su3_mult(cplx res[2][3], const double* u, cplx h[2][3]) {
  for (int s = 0; s < 2; ++s)
    for (int row = 0; row < 3; ++row) {
      cplx acc = {0.0, 0.0};
      for (int col = 0; col < 3; ++col) {
        cplx m = { u[(col*3+row)*2], u[(col*3+row)*2+1] };
        acc = c_add(acc, c_mul(m, h[s][col]));
      }
      res[s][row] = acc;
    }
}
static inline void su3_adj_mult(cplx res[2][3], const double* u, cplx h[2][3]) {
  for (int s = 0; s < 2; ++s)
    for (int row = 0; row < 3; ++row) {
      cplx acc = {0.0, 0.0};
      for (int col = 0; col < 3; ++col) {
        cplx m = { u[(row*3+col)*2], -u[(row*3+col)*2+1] };
        acc = c_add(acc, c_mul(m, h[s][col]));
      }
      res[s][row] = acc;
    }
}

/* Spin projection (1 + sgn*gamma_mu) onto the upper two components and the matching
 * reconstruction of the lower two, DeGrand-Rossi basis, restated from
 * cpp_dslash_scalar_64bit_c.h:33-1190 (one inline function per direction/sign there). */
static inline void project(cplx h[2][3], const double* a, int mu, int sgn) {
  const cplx* A = (const cplx*)a; /* A[spin*3+colour] */
  for (int c = 0; c < 3; ++c) {
    cplx a0 = A[c], a1 = A[3+c], a2 = A[6+c], a3 = A[9+c];
    switch (mu) {
      case 0: /* (1-g0): a0 - i a3, a1 - i a2 */
        if (sgn < 0) { h[0][c] = c_sub(a0, c_timesI(a3)); h[1][c] = c_sub(a1, c_timesI(a2)); }
        else         { h[0][c] = c_add(a0, c_timesI(a3)); h[1][c] = c_add(a1, c_timesI(a2)); }
        break;
      case 1: /* (1-g1): a0 + a3, a1 - a2 */
        if (sgn < 0) { h[0][c] = c_add(a0, a3); h[1][c] = c_sub(a1, a2); }
        else         { h[0][c] = c_sub(a0, a3); h[1][c] = c_add(a1, a2); }
        break;
      case 2: /* (1-g2): a0 - i a2, a1 + i a3 */
        if (sgn < 0) { h[0][c] = c_sub(a0, c_timesI(a2)); h[1][c] = c_add(a1, c_timesI(a3)); }
        else         { h[0][c] = c_add(a0, c_timesI(a2)); h[1][c] = c_sub(a1, c_timesI(a3)); }
        break;
      default: /* (1-g3): a0 - a2, a1 - a3 */
        if (sgn < 0) { h[0][c] = c_sub(a0, a2); h[1][c] = c_sub(a1, a3); }
        else         { h[0][c] = c_add(a0, a2); h[1][c] = c_add(a1, a3); }
        break;
    }
  }
}
static inline void recons_add(cplx* out, cplx r[2][3], int mu, int sgn) {
  for (int c = 0; c < 3; ++c) {
    cplx r0 = r[0][c], r1 = r[1][c], r2, r3;
    switch (mu) {
      case 0: if (sgn < 0) { r2 = c_timesI(r1); r3 = c_timesI(r0); }
              else { r2 = c_neg(c_timesI(r1)); r3 = c_neg(c_timesI(r0)); } break;
      case 1: if (sgn < 0) { r2 = c_neg(r1); r3 = r0; }
              else { r2 = r1; r3 = c_neg(r0); } break;
      case 2: if (sgn < 0) { r2 = c_timesI(r0); r3 = c_neg(c_timesI(r1)); }
              else { r2 = c_neg(c_timesI(r0)); r3 = c_timesI(r1); } break;
      default: if (sgn < 0) { r2 = c_neg(r0); r3 = c_neg(r1); }
              else { r2 = r0; r3 = r1; } break;
    }
    out[c]   = c_add(out[c],   r0);
    out[3+c] = c_add(out[3+c], r1);
    out[6+c] = c_add(out[6+c], r2);
    out[9+c] = c_add(out[9+c], r3);
  }
}

typedef struct {
  int L[4]; int V, Vh;
  int *fwd, *bwd;
} orc_geom;

orc_geom* orc_geom_create(const int L[4]) {
  orc_geom* g = (orc_geom*)calloc(1, sizeof(orc_geom));
  for (int i = 0; i < 4; ++i) g->L[i] = L[i];
  g->V = L[0]*L[1]*L[2]*L[3]; g->Vh = g->V/2;
  build_tables(L, &g->fwd, &g->bwd);
  return g;
}
void orc_geom_free(orc_geom* g) { if (g) { free(g->fwd); free(g->bwd); free(g); } }

/* Dslash<double>::operator()(res, psi, u, isign, cb): cb is the SOURCE checkerboard, the
 * loop runs over target sites of parity 1-cb (cpp_dslash_scalar_64bit.cc:35-65, 69-214;
 * lwldslash_w_cppd.cc:196-205).  isign=+1: forward hop (1-g_mu) U_mu(x) psi(x+mu),
 * backward hop (1+g_mu) U_mu(x-mu)^dag psi(x-mu); isign=-1 swaps the projector signs
 * (lwldslash_w.h:252-295). */
void orc_dslash(const orc_geom* g, double* res, const double* psi, const double* packed_u, int isign, int cb) {
  int Vh = g->Vh, tcb = 1 - cb;
#pragma omp parallel for
  for (int i = 0; i < Vh; ++i) {
    int ix = tcb*Vh + i;
    cplx acc[12]; memset(acc, 0, sizeof(acc));
    for (int mu = 0; mu < 4; ++mu) {
      cplx h[2][3], r[2][3];
      int f = g->fwd[4*ix+mu], b = g->bwd[4*ix+mu];
      project(h, psi + (size_t)f*SPINOR, mu, -isign);
      su3_mult(r, packed_u + ((size_t)ix*4+mu)*LINK, h);
      recons_add(acc, r, mu, -isign);
      project(h, psi + (size_t)b*SPINOR, mu, +isign);
      su3_adj_mult(r, packed_u + ((size_t)b*4+mu)*LINK, h);
      recons_add(acc, r, mu, +isign);
    }
    memcpy(res + (size_t)ix*SPINOR, acc, sizeof(acc));
  }
}

/* ------------------------------------------------------------------------- */
/* Field strength: lib/meas/glue/mesfield.cc:44-74                             */
/* ------------------------------------------------------------------------- */
static inline void m_mul(cplx* r, const cplx* a, const cplx* b) {          /* r = a b */
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
    cplx s = {0,0};
    for (int k = 0; k < 3; ++k) s = c_add(s, c_mul(a[i*3+k], b[k*3+j]));
    r[i*3+j] = s;
  }
}
static inline void m_mul_adj(cplx* r, const cplx* a, const cplx* b) {      /* r = a b^dag */
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
    cplx s = {0,0};
    for (int k = 0; k < 3; ++k) s = c_add(s, c_mul(a[i*3+k], c_conj(b[j*3+k])));
    r[i*3+j] = s;
  }
}
static inline void m_adj_mul(cplx* r, const cplx* a, const cplx* b) {      /* r = a^dag b */
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
    cplx s = {0,0};
    for (int k = 0; k < 3; ++k) s = c_add(s, c_mul(c_conj(a[k*3+i]), b[k*3+j]));
    r[i*3+j] = s;
  }
}

/* f[offset] for (mu<nu) in the order (0,1),(0,2),(0,3),(1,2),(1,3),(2,3); f is [6][V][3][3][2]. */
void orc_mesfield(const orc_geom* g, const double* const u[4], double* f) {
  int V = g->V;
  cplx* t2 = (cplx*)malloc(sizeof(cplx)*9*(size_t)V);  /* adj(tmp_0)*tmp_1            */
  cplx* t3 = (cplx*)malloc(sizeof(cplx)*9*(size_t)V);  /* tmp_0*adj(tmp_1), 2nd stage  */
  cplx* t4 = (cplx*)malloc(sizeof(cplx)*9*(size_t)V);  /* adj(tmp_1)*tmp_0, 2nd stage  */
  int offset = 0;
  for (int mu = 0; mu < 3; ++mu)
    for (int nu = mu+1; nu < 4; ++nu, ++offset) {
      cplx* F = (cplx*)(f + (size_t)offset*V*LINK);
      const cplx* Um = (const cplx*)u[mu];
      const cplx* Un = (const cplx*)u[nu];
#pragma omp parallel for
      for (int s = 0; s < V; ++s) {
        const cplx* tmp_3 = Un + 9*(size_t)g->fwd[4*s+mu];   /* shift(u[nu],FORWARD,mu) */
        const cplx* tmp_4 = Um + 9*(size_t)g->fwd[4*s+nu];   /* shift(u[mu],FORWARD,nu) */
        cplx tmp_0[9], tmp_1[9];
        m_mul(tmp_0, Un + 9*(size_t)s, tmp_4);               /* u[nu]*tmp_4 */
        m_mul(tmp_1, Um + 9*(size_t)s, tmp_3);               /* u[mu]*tmp_3 */
        m_mul_adj(F + 9*(size_t)s, tmp_1, tmp_0);            /* f = tmp_1*adj(tmp_0) */
        m_adj_mul(t2 + 9*(size_t)s, tmp_0, tmp_1);           /* tmp_2 = adj(tmp_0)*tmp_1 */
        cplx a[9], b[9];
        m_mul_adj(a, tmp_4, tmp_3);                          /* tmp_1 = tmp_4*adj(tmp_3) */
        m_adj_mul(b, Un + 9*(size_t)s, Um + 9*(size_t)s);    /* tmp_0 = adj(u[nu])*u[mu] */
        m_mul_adj(t3 + 9*(size_t)s, b, a);                   /* tmp_0*adj(tmp_1) */
        m_adj_mul(t4 + 9*(size_t)s, a, b);                   /* adj(tmp_1)*tmp_0 */
      }
#pragma omp parallel for
      for (int s = 0; s < V; ++s) {
        int sbn = g->bwd[4*s+nu], sbm = g->bwd[4*s+mu];
        int sbnm = g->bwd[4*sbn+mu];                          /* shift(shift(.,BACKWARD,nu),BACKWARD,mu) */
        cplx* Fs = F + 9*(size_t)s;
        for (int k = 0; k < 9; ++k) {
          Fs[k] = c_add(Fs[k], t2[9*(size_t)sbnm+k]);
          Fs[k] = c_add(Fs[k], t3[9*(size_t)sbn+k]);
          Fs[k] = c_add(Fs[k], t4[9*(size_t)sbm+k]);
        }
        cplx adjF[9];
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) adjF[i*3+j] = c_conj(Fs[j*3+i]);
        for (int k = 0; k < 9; ++k) { Fs[k] = c_sub(Fs[k], adjF[k]); Fs[k].re *= 0.125; Fs[k].im *= 0.125; }
      }
    }
  free(t2); free(t3); free(t4);
}

/* ------------------------------------------------------------------------- */
/* Clover term: coefficients, makeClov, LDL^dag inverse, apply                 */
/* ------------------------------------------------------------------------- */
/* QDPCloverTermT::create, clover_term_qdp_w.h:263-278: clovCoeffR *= 0.5/xi_0 (aniso) or 0.5,
 * clovCoeffT *= 0.5, diag_mass = 1 + (Nd-1)*(nu/xi_0 | 1) + Mass. out = {diag_mass, cR, cT}. */
void orc_clover_coeffs(double Mass, double clovCoeffR, double clovCoeffT, int anisoP, double xi_0, double nu, double out[3]) {
  double ff = anisoP ? 1.0/xi_0 : 1.0;
  out[1] = clovCoeffR * 0.5 * ff;
  out[2] = clovCoeffT * 0.5;
  double fm = anisoP ? nu/xi_0 : 1.0;
  out[0] = 1.0 + 3.0*fm + Mass;
}

/* makeFermCoeffs, lib/io/aniso_io.cc:63-80 */
void orc_ferm_coeffs(int anisoP, int t_dir, double xi_0, double nu, double out[4]) {
  for (int mu = 0; mu < 4; ++mu) out[mu] = (anisoP && mu != t_dir) ? nu/xi_0 : 1.0;
}

/* makeClov + makeClovSiteLoop, clover_term_qdp_w.h:398-553; getCloverCoeff :1524-1544.
 * tri is [V][72]: diag[0][0..5], diag[1][0..5], offd[0][0..14][2], offd[1][0..14][2]. */
void orc_make_clov(const orc_geom* g, const double* f, double diag_mass, double cR, double cT,
                   int anisoP, int t_dir, double* tri) {
  int V = g->V;
  static const int pmu[6] = {0,0,0,1,1,2}, pnu[6] = {1,2,3,2,3,3};
  double coef[6];
  for (int k = 0; k < 6; ++k)
    coef[k] = (anisoP && (pmu[k] == t_dir || pnu[k] == t_dir)) ? cT : cR;
#pragma omp parallel for
  for (int site = 0; site < V; ++site) {
    cplx F[6][9];
    for (int k = 0; k < 6; ++k) {
      const cplx* fk = (const cplx*)(f + ((size_t)k*V + site)*LINK);
      for (int e = 0; e < 9; ++e) { F[k][e].re = fk[e].re*coef[k]; F[k][e].im = fk[e].im*coef[k]; }
    }
    double* diag0 = tri + (size_t)site*CLOV;
    double* diag1 = diag0 + 6;
    cplx* offd0 = (cplx*)(diag0 + 12);
    cplx* offd1 = offd0 + 15;
    for (int i = 0; i < 6; ++i) { diag0[i] = diag_mass; diag1[i] = diag_mass; }
    for (int i = 0; i < NC; ++i) {
      cplx c0 = c_sub(F[5][i*3+i], F[0][i*3+i]);
      diag0[i] += c0.im; diag0[i+NC] -= c0.im;
      cplx c1 = c_add(F[5][i*3+i], F[0][i*3+i]);
      diag1[i] -= c1.im; diag1[i+NC] += c1.im;
    }
    for (int i = 1; i < NC; ++i)
      for (int j = 0; j < i; ++j) {
        int eij = i*(i-1)/2 + j, etmp = (i+NC)*(i+NC-1)/2 + j + NC;
        offd0[eij] = c_timesI(c_sub(F[0][i*3+j], F[5][i*3+j]));
        offd0[etmp] = c_neg(offd0[eij]);
        offd1[eij] = c_timesI(c_add(F[5][i*3+j], F[0][i*3+j]));
        offd1[etmp] = c_neg(offd1[eij]);
      }
    for (int i = 0; i < NC; ++i)
      for (int j = 0; j < NC; ++j) {
        int eij = (i+NC)*(i+NC-1)/2 + j;
        cplx E_minus = c_add(c_timesI(F[2][i*3+j]), F[4][i*3+j]);
        cplx B_minus = c_sub(c_timesI(F[3][i*3+j]), F[1][i*3+j]);
        offd0[eij] = c_sub(B_minus, E_minus);
        offd1[eij] = c_add(E_minus, B_minus);
      }
  }
}

/* ldagdlinv / LDagDLInvSiteLoop, clover_term_qdp_w.h:619-846: in-place LDL^dag factorisation
 * (Golub & van Loan alg. 4.1.2) and inversion of both 6x6 blocks on checkerboard cb.
 * tr_log_diag (length V, may be NULL) receives sum log|d_i| on the sites touched. */
void orc_ldagdlinv(const orc_geom* g, double* tri, int cb, double* tr_log_diag) {
  int Vh = g->Vh;
  const int N = 6;
#pragma omp parallel for
  for (int ss = 0; ss < Vh; ++ss) {
    int site = cb*Vh + ss;
    double tl = 0.0;
    for (int block = 0; block < 2; ++block) {
      double* d_ptr = tri + (size_t)site*CLOV + 6*block;
      cplx* o_ptr = (cplx*)(tri + (size_t)site*CLOV + 12) + 15*block;
      double inv_d[6], diag_g[6]; cplx inv_offd[15], v[6];
      for (int i = 0; i < N; ++i) inv_d[i] = d_ptr[i];
      for (int i = 0; i < 15; ++i) inv_offd[i] = o_ptr[i];
      for (int j = 0; j < N; ++j) {
        for (int i = 0; i < j; ++i) {
          int eji = j*(j-1)/2 + i;
          cplx Aii = { inv_d[i], 0.0 };
          v[i] = c_mul(Aii, c_conj(inv_offd[eji]));
        }
        v[j].re = inv_d[j]; v[j].im = 0.0;
        for (int k = 0; k < j; ++k) {
          int ejk = j*(j-1)/2 + k;
          v[j] = c_sub(v[j], c_mul(inv_offd[ejk], v[k]));
        }
        inv_d[j] = v[j].re;
        for (int k = j+1; k < N; ++k) {
          int ekj = k*(k-1)/2 + j;
          for (int l = 0; l < j; ++l) {
            int ekl = k*(k-1)/2 + l;
            inv_offd[ekj] = c_sub(inv_offd[ekj], c_mul(inv_offd[ekl], v[l]));
          }
          inv_offd[ekj] = c_div(inv_offd[ekj], v[j]);
        }
      }
      for (int i = 0; i < N; ++i) { diag_g[i] = 1.0/inv_d[i]; tl += log(fabs(inv_d[i])); }
      for (int k = 0; k < N; ++k) {
        for (int i = 0; i < k; ++i) { v[i].re = 0; v[i].im = 0; }
        v[k].re = diag_g[k]; v[k].im = 0.0;
        for (int i = k+1; i < N; ++i) {
          v[i].re = 0; v[i].im = 0;
          for (int j = k; j < i; ++j) {
            int eij = i*(i-1)/2 + j;
            cplx dj = { inv_d[j], 0.0 };
            v[i] = c_sub(v[i], c_mul(c_mul(inv_offd[eij], dj), v[j]));
          }
          v[i].re *= diag_g[i]; v[i].im *= diag_g[i];
        }
        for (int i = N-2; i >= k; --i)
          for (int j = i+1; j < N; ++j) {
            int eji = j*(j-1)/2 + i;
            v[i] = c_sub(v[i], c_mul(c_conj(inv_offd[eji]), v[j]));
          }
        inv_d[k] = v[k].re;
        for (int i = k+1; i < N; ++i) inv_offd[i*(i-1)/2 + k] = v[i];
      }
      for (int i = 0; i < N; ++i) d_ptr[i] = inv_d[i];
      for (int i = 0; i < 15; ++i) o_ptr[i] = inv_offd[i];
    }
    if (tr_log_diag) tr_log_diag[site] = tl;
  }
}

/* applySiteLoop, clover_term_qdp_w.h:1562-1634: chi = (L + D + L^dag) psi on checkerboard cb. */
static inline void clover_site_apply(double* chi, const double* psi, const double* tri) {
  const int n = 6;
  cplx* cchi = (cplx*)chi; const cplx* ppsi = (const cplx*)psi;
  const double* diag0 = tri; const double* diag1 = tri + 6;
  const cplx* off0 = (const cplx*)(tri + 12); const cplx* off1 = off0 + 15;
  cplx out[12];
  for (int i = 0; i < n; ++i) {
    out[i].re = diag0[i]*ppsi[i].re;     out[i].im = diag0[i]*ppsi[i].im;
    out[n+i].re = diag1[i]*ppsi[n+i].re; out[n+i].im = diag1[i]*ppsi[n+i].im;
  }
  int kij = 0;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < i; ++j) {
      out[i]   = c_add(out[i],   c_mul(off0[kij], ppsi[j]));
      out[j]   = c_add(out[j],   c_mul(c_conj(off0[kij]), ppsi[i]));
      out[n+i] = c_add(out[n+i], c_mul(off1[kij], ppsi[n+j]));
      out[n+j] = c_add(out[n+j], c_mul(c_conj(off1[kij]), ppsi[n+i]));
      ++kij;
    }
  memcpy(cchi, out, sizeof(out));
}

void orc_clover_apply(const orc_geom* g, double* chi, const double* psi, const double* tri, int cb) {
  int Vh = g->Vh;
#pragma omp parallel for
  for (int ss = 0; ss < Vh; ++ss) {
    size_t site = (size_t)cb*Vh + ss;
    clover_site_apply(chi + site*SPINOR, psi + site*SPINOR, tri + site*CLOV);
  }
}

/* ------------------------------------------------------------------------- */
/* EvenOddPrecCloverLinOp                                                      */
/* ------------------------------------------------------------------------- */
typedef struct {
  orc_geom* g;
  double* packed_u;   /* aniso-folded, transposed links */
  double* clov;       /* A on both checkerboards */
  double* invclov;    /* copy of A with cb 0 inverted (A_ee^-1) */
  double* tmp1; double* tmp2;
  double* tr_log_diag;
  int sym;            /* 0: EvenOddPrecCloverLinOp, 1: SymEvenOddPrecCloverLinOp (invclov cb 1 inverted too) */
  double twisted_m;   /* CloverFermActParams::twisted_m; 0 = twisted_m_usedP false */
} orc_op;

/* EvenOddPrecCloverLinOp::create, eoprec_clover_linop_w.cc:19-39: clov.create (mesField +
 * makeClov on the BC-modified links), invclov = copy, invclov.choles(0), D.create (aniso
 * factors folded into the links, then packed).  u[] must already carry the fermion BC phases
 * (state->getLinks()). */
orc_op* orc_op_create(const int L[4], const double* const u[4], double Mass, double clovCoeffR,
                      double clovCoeffT, int anisoP, int t_dir, double xi_0, double nu) {
  orc_op* op = (orc_op*)calloc(1, sizeof(orc_op));
  op->g = orc_geom_create(L);
  size_t V = (size_t)op->g->V;
  double cc[3], fc[4];
  orc_clover_coeffs(Mass, clovCoeffR, clovCoeffT, anisoP, xi_0, nu, cc);
  orc_ferm_coeffs(anisoP, t_dir, xi_0, nu, fc);
  op->packed_u = (double*)malloc(sizeof(double)*V*4*LINK);
  orc_pack_gauge(L, u, fc, op->packed_u);
  double* f = (double*)malloc(sizeof(double)*6*V*LINK);
  orc_mesfield(op->g, u, f);
  op->clov = (double*)malloc(sizeof(double)*V*CLOV);
  orc_make_clov(op->g, f, cc[0], cc[1], cc[2], anisoP, t_dir, op->clov);
  free(f);
  op->invclov = (double*)malloc(sizeof(double)*V*CLOV);
  memcpy(op->invclov, op->clov, sizeof(double)*V*CLOV);
  op->tr_log_diag = (double*)calloc(V, sizeof(double));
  orc_ldagdlinv(op->g, op->invclov, 0, op->tr_log_diag);
  op->tmp1 = (double*)calloc(V*SPINOR, sizeof(double));
  op->tmp2 = (double*)calloc(V*SPINOR, sizeof(double));
  return op;
}
void orc_op_free(orc_op* op) {
  if (!op) return;
  orc_geom_free(op->g); free(op->packed_u); free(op->clov); free(op->invclov);
  free(op->tmp1); free(op->tmp2); free(op->tr_log_diag); free(op);
}
const double* orc_op_clov(const orc_op* op) { return op->clov; }
const double* orc_op_invclov(const orc_op* op) { return op->invclov; }
const double* orc_op_packed_gauge(const orc_op* op) { return op->packed_u; }
const orc_geom* orc_op_geom(const orc_op* op) { return op->g; }

/* D.apply(chi, psi, isign, cb): cb is the TARGET checkerboard here, the library call gets
 * source_cb = 1-cb (lwldslash_w_cppd.cc:196-205). */
void orc_op_dslash(const orc_op* op, double* chi, const double* psi, int isign, int cb) {
  orc_dslash(op->g, chi, psi, op->packed_u, isign, 1 - cb);
}

/* The twisted-mass term both operators end with (eoprec_clover_linop_w.cc:174-184, seoprec_clover_linop_w.cc:174-184):
 * tmp1 = Gamma(15) * timesI(psi); chi += twisted_m * tmp1 (PLUS) / chi -= twisted_m * tmp1 (MINUS), odd checkerboard.
 * Gamma(15) = gamma_0 gamma_1 gamma_2 gamma_3 = diag(1,1,-1,-1) in the basis of the projectors above. */
void orc_op_set_twisted_mass(orc_op* op, double mu) { op->twisted_m = mu; }
static void orc_twisted_term(const orc_op* op, double* chi, const double* psi, int isign) {
  if (op->twisted_m == 0.0) return;
  int Vh = op->g->Vh;
  const double mu = (isign > 0 ? 1.0 : -1.0) * op->twisted_m;
#pragma omp parallel for
  for (int i = Vh; i < 2*Vh; ++i) {
    for (int s = 0; s < 4; ++s) {
      const double g5 = s < 2 ? 1.0 : -1.0;
      for (int c = 0; c < 3; ++c) {
        const double* p = psi + (size_t)i*SPINOR + (s*3 + c)*2;
        double* x = chi + (size_t)i*SPINOR + (s*3 + c)*2;
        const double tr = g5 * (-p[1]), ti = g5 * p[0];     /* Gamma(15) * (i psi) */
        x[0] += mu * tr; x[1] += mu * ti;
      }
    }
  }
}

/* EvenOddPrecCloverLinOp::operator(), eoprec_clover_linop_w.cc:142-187:
 * tmp1 = D_eo psi; tmp2 = A_ee^-1 tmp1; tmp1 = D_oe tmp2; chi = A_oo psi; chi -= 1/4 tmp1.
 * Fields are full-lattice arrays; only the odd half of chi is written. */
static void orc_sym_apply(orc_op* op, double* chi, const double* psi, int isign);
void orc_op_apply(orc_op* op, double* chi, const double* psi, int isign) {
  int Vh = op->g->Vh;
  if (op->sym) { orc_sym_apply(op, chi, psi, isign); return; }
  orc_op_dslash(op, op->tmp1, psi, isign, 0);
  orc_clover_apply(op->g, op->tmp2, op->tmp1, op->invclov, 0);
  orc_op_dslash(op, op->tmp1, op->tmp2, isign, 1);
  orc_clover_apply(op->g, chi, psi, op->clov, 1);
  double* c = chi + (size_t)Vh*SPINOR; const double* t = op->tmp1 + (size_t)Vh*SPINOR;
  size_t n = (size_t)Vh*SPINOR;
#pragma omp parallel for
  for (size_t i = 0; i < n; ++i) c[i] += -0.25*t[i];
  orc_twisted_term(op, chi, psi, isign);
}

/* ------------------------------------------------------------------------- */
/* SymEvenOddPrecCloverLinOp  (section 8 f4)                                   */
/* ------------------------------------------------------------------------- */
/* SymEvenOddPrecCloverLinOp::create, seoprec_clover_linop_w.cc:16-41: as the asymmetric create plus
 * invclov.choles(1).  Switching it on inverts the cb-1 half of invclov in place (once); tr_log_diag then holds
 * log|det| on both checkerboards (logDetEvenEvenLinOp + logDetOddOddLinOp of the symmetric log-det operator). */
void orc_op_set_symmetric(orc_op* op, int sym) {
  if (sym && !op->sym) orc_ldagdlinv(op->g, op->invclov, 1, op->tr_log_diag);
  if (!sym && op->sym) {   /* restore A_oo in invclov so that a later switch inverts it again */
    size_t n = (size_t)op->g->Vh*CLOV;
    memcpy(op->invclov + n, op->clov + n, n*sizeof(double));
  }
  op->sym = sym ? 1 : 0;
}
int orc_op_is_symmetric(const orc_op* op) { return op->sym; }
double orc_op_tr_log(const orc_op* op, int cb) {
  double s = 0; int Vh = op->g->Vh;
  for (int i = 0; i < Vh; ++i) s += op->tr_log_diag[cb*Vh + i];
  return s;
}

/* SymEvenOddPrecCloverLinOp::operator(), seoprec_clover_linop_w.cc:147-193 (no twisted mass):
 * PLUS : tmp1 = D_eo psi; tmp2 = A_ee^-1 tmp1; tmp1 = D_oe tmp2; tmp2 = A_oo^-1 tmp1
 * MINUS: tmp1 = A_oo^-1 psi; tmp2 = D_eo^dag tmp1; tmp1 = A_ee^-1 tmp2; tmp2 = D_oe^dag tmp1
 * chi = psi - 1/4 tmp2 on rb[1]. */
static void orc_sym_apply(orc_op* op, double* chi, const double* psi, int isign) {
  int Vh = op->g->Vh; size_t n = (size_t)Vh*SPINOR;
  if (isign > 0) {
    orc_op_dslash(op, op->tmp1, psi, isign, 0);
    orc_clover_apply(op->g, op->tmp2, op->tmp1, op->invclov, 0);
    orc_op_dslash(op, op->tmp1, op->tmp2, isign, 1);
    orc_clover_apply(op->g, op->tmp2, op->tmp1, op->invclov, 1);
  } else {
    orc_clover_apply(op->g, op->tmp1, psi, op->invclov, 1);
    orc_op_dslash(op, op->tmp2, op->tmp1, isign, 0);
    orc_clover_apply(op->g, op->tmp1, op->tmp2, op->invclov, 0);
    orc_op_dslash(op, op->tmp2, op->tmp1, isign, 1);
  }
  const double* t = op->tmp2 + n; const double* x = psi + n; double* c = chi + n;
#pragma omp parallel for
  for (size_t i = 0; i < n; ++i) c[i] = x[i] + -0.25*t[i];
  orc_twisted_term(op, chi, psi, isign);
}

/* ------------------------------------------------------------------------- */
/* Subset BLAS on rb[1] (odd half of a full-lattice array)                     */
/* ------------------------------------------------------------------------- */
static double norm2_odd(const orc_geom* g, const double* x) {
  size_t n = (size_t)g->Vh*SPINOR; const double* p = x + n; double s = 0;
#pragma omp parallel for reduction(+:s)
  for (size_t i = 0; i < n; ++i) s += p[i]*p[i];
  return s;
}
static cplx inner_odd(const orc_geom* g, const double* x, const double* y) {  /* <x|y> = sum conj(x) y */
  size_t n = (size_t)g->Vh*12; const cplx* a = (const cplx*)x + n; const cplx* b = (const cplx*)y + n;
  double sr = 0, si = 0;
#pragma omp parallel for reduction(+:sr,si)
  for (size_t i = 0; i < n; ++i) { sr += a[i].re*b[i].re + a[i].im*b[i].im; si += a[i].re*b[i].im - a[i].im*b[i].re; }
  cplx r = { sr, si }; return r;
}
double orc_norm2_odd(const orc_geom* g, const double* x) { return norm2_odd(g, x); }
void orc_inner_odd(const orc_geom* g, const double* x, const double* y, double out[2]) {
  cplx r = inner_odd(g, x, y); out[0] = r.re; out[1] = r.im;
}

/* ------------------------------------------------------------------------- */
/* InvCG2_a, lib/actions/ferm/invert/invcg2.cc:70-232                          */
/* ------------------------------------------------------------------------- */
/* Solves (M^dag M) psi = chi on rb[1]. Returns n_count; *resid as the reference sets it.
 * max_iter_timing: if >0, never tests convergence (used for fixed-iteration CPU timing). */
int orc_invcg2(orc_op* op, const double* chi, double* psi, double RsdCG, int MaxCG, double* resid) {
  const orc_geom* g = op->g;
  size_t V = (size_t)g->V, n = (size_t)g->Vh*SPINOR, off = n;
  double* mp = (double*)calloc(V*SPINOR, sizeof(double));
  double* mmp = (double*)calloc(V*SPINOR, sizeof(double));
  double* p = (double*)calloc(V*SPINOR, sizeof(double));
  double* r = (double*)calloc(V*SPINOR, sizeof(double));
  int n_count = MaxCG;
  double chi_sq = norm2_odd(g, chi);
  double rsd_sq = (RsdCG*RsdCG)*chi_sq;
  orc_op_apply(op, mp, psi, +1);
  orc_op_apply(op, mmp, mp, -1);
#pragma omp parallel for
  for (size_t i = 0; i < n; ++i) r[off+i] = chi[off+i] - mmp[off+i];
  double cp = norm2_odd(g, r);
  memcpy(p + off, r + off, n*sizeof(double));
  if (cp <= rsd_sq) { *resid = sqrt(cp); n_count = 0; goto done; }
  {
    double a, b, c, d;
    int k;
    for (k = 1; k <= MaxCG; ++k) {
      c = cp;
      orc_op_apply(op, mp, p, +1);
      d = norm2_odd(g, mp);
      orc_op_apply(op, mmp, mp, -1);
      a = c/d;
#pragma omp parallel for
      for (size_t i = 0; i < n; ++i) r[off+i] -= a*mmp[off+i];
      cp = norm2_odd(g, r);
#pragma omp parallel for
      for (size_t i = 0; i < n; ++i) psi[off+i] += a*p[off+i];
      if (cp <= rsd_sq) {
        n_count = k;
        orc_op_apply(op, mp, psi, +1);
        orc_op_apply(op, mmp, mp, -1);
        double s = 0;
#pragma omp parallel for reduction(+:s)
        for (size_t i = 0; i < n; ++i) { double t = chi[off+i] - mmp[off+i]; s += t*t; }
        *resid = sqrt(s);
        goto done;
      }
      b = cp/c;
#pragma omp parallel for
      for (size_t i = 0; i < n; ++i) p[off+i] = r[off+i] + b*p[off+i];
    }
    n_count = MaxCG; *resid = sqrt(cp);
  }
done:
  free(mp); free(mmp); free(p); free(r);
  return n_count;
}

/* LinOpSysSolverCG::operator(), syssolver_linop_cg.h:57-96: chi_tmp = M^dag chi; InvCG2;
 * resid = |chi - M psi|.  Returns n_count; out[0] = resid, out[1] = relative resid. */
int orc_solve_cg(orc_op* op, const double* chi, double* psi, double RsdCG, int MaxCG, double out[2]) {
  const orc_geom* g = op->g;
  size_t V = (size_t)g->V, n = (size_t)g->Vh*SPINOR, off = n;
  double* chi_tmp = (double*)calloc(V*SPINOR, sizeof(double));
  double dummy;
  orc_op_apply(op, chi_tmp, chi, -1);
  int n_count = orc_invcg2(op, chi_tmp, psi, RsdCG, MaxCG, &dummy);
  orc_op_apply(op, chi_tmp, psi, +1);
  double s = 0;
#pragma omp parallel for reduction(+:s)
  for (size_t i = 0; i < n; ++i) { double t = chi[off+i] - chi_tmp[off+i]; s += t*t; }
  out[0] = sqrt(s); out[1] = out[0]/sqrt(norm2_odd(g, chi));
  free(chi_tmp);
  return n_count;
}

/* ------------------------------------------------------------------------- */
/* MInvCG2_a, lib/actions/ferm/invert/minvcg2.cc:74-373  (section 8 f4)       */
/* ------------------------------------------------------------------------- */
/* Multi-shift CG: (M^dag M + shifts[s]) psi[s] = chi on rb[1] for all s at once.  psi = n_shift full-lattice
 * arrays back to back (all zeroed on entry, as the reference does).  Returns n_count. */
int orc_minvcg2(orc_op* op, const double* chi, double* psi, const double* shifts, const double* RsdCG,
                int n_shift, int MaxCG) {
  const orc_geom* g = op->g;
  size_t V = (size_t)g->V, n = (size_t)g->Vh*SPINOR, off = n, FS = V*SPINOR;
  int isz = 0, n_count = 0, s, k;
  for (s = 1; s < n_shift; ++s) if (shifts[s] < shifts[isz]) isz = s;
  memset(psi, 0, sizeof(double)*FS*n_shift);
  double chi_norm_sq = norm2_odd(g, chi);
  if (sqrt(chi_norm_sq) < 1.0e-5) return 0;                               /* fuzz, minvcg2.cc:135-148 */
  double* rsd_sq = (double*)malloc(sizeof(double)*n_shift);
  double cp = chi_norm_sq;
  for (s = 0; s < n_shift; ++s) rsd_sq[s] = cp*RsdCG[s]*RsdCG[s];
  double* r = (double*)calloc(FS, sizeof(double));
  double* p0 = (double*)calloc(FS, sizeof(double));
  double* p = (double*)calloc(FS*n_shift, sizeof(double));
  double* Mp = (double*)calloc(FS, sizeof(double));
  double* MMp = (double*)calloc(FS, sizeof(double));
  memcpy(r + off, chi + off, n*sizeof(double));
  memcpy(p0 + off, chi + off, n*sizeof(double));
  for (s = 0; s < n_shift; ++s) memcpy(p + s*FS + off, chi + off, n*sizeof(double));
  orc_op_apply(op, Mp, p0, +1);
  double d = norm2_odd(g, Mp);
  orc_op_apply(op, MMp, Mp, -1);
  double b = -cp/d;
#pragma omp parallel for
  for (size_t i = 0; i < n; ++i) r[off+i] += b*MMp[off+i];
  double* bs = (double*)malloc(sizeof(double)*n_shift);
  double* z[2]; z[0] = (double*)malloc(sizeof(double)*n_shift); z[1] = (double*)malloc(sizeof(double)*n_shift);
  int* convsP = (int*)calloc(n_shift, sizeof(int));
  int iz = 1;
  for (s = 0; s < n_shift; ++s) {
    z[1-iz][s] = 1.0;
    z[iz][s] = 1.0/(1.0 - shifts[s]*b);
    bs[s] = b*z[iz][s];
  }
  for (s = 0; s < n_shift; ++s) {
    double* ps = psi + s*FS; double f = bs[s];
#pragma omp parallel for
    for (size_t i = 0; i < n; ++i) ps[off+i] = -f*chi[off+i];
  }
  double c = norm2_odd(g, r);
  int convP = c < rsd_sq[isz];
  for (k = 1; k <= MaxCG && !convP; ++k) {
    double a = c/cp;
#pragma omp parallel for
    for (size_t i = 0; i < n; ++i) p0[off+i] = r[off+i] + a*p0[off+i];
    for (s = 0; s < n_shift; ++s) if (!convsP[s]) {
      double as = a*z[iz][s]*bs[s]/(z[1-iz][s]*b), zz = z[iz][s];
      double* pp = p + s*FS;
#pragma omp parallel for
      for (size_t i = 0; i < n; ++i) pp[off+i] = zz*r[off+i] + as*pp[off+i];
    }
    cp = c;
    orc_op_apply(op, Mp, p0, +1);
    d = norm2_odd(g, Mp);
    orc_op_apply(op, MMp, Mp, -1);
    double bp = b;
    b = -cp/d;
#pragma omp parallel for
    for (size_t i = 0; i < n; ++i) r[off+i] += b*MMp[off+i];
    c = norm2_odd(g, r);
    iz = 1 - iz;
    for (s = 0; s < n_shift; ++s) if (!convsP[s]) {
      double z0 = z[1-iz][s], z1 = z[iz][s];
      z[iz][s] = z0*z1*bp;
      z[iz][s] /= b*a*(z1 - z0) + z1*bp*(1.0 - shifts[s]*b);
      bs[s] = b*z[iz][s]/z0;
    }
    for (s = 0; s < n_shift; ++s) if (!convsP[s]) {
      double f = bs[s]; double* ps = psi + s*FS; const double* pp = p + s*FS;
#pragma omp parallel for
      for (size_t i = 0; i < n; ++i) ps[off+i] -= f*pp[off+i];
    }
    convP = 1;
    for (s = 0; s < n_shift; ++s) {
      if (!convsP[s]) {
        double css = c*z[iz][s]*z[iz][s];
        convsP[s] = css < rsd_sq[s];
      }
      convP &= convsP[s];
    }
    n_count = k;
  }
  free(rsd_sq); free(r); free(p0); free(p); free(Mp); free(MMp); free(bs); free(z[0]); free(z[1]); free(convsP);
  return n_count;
}

/* ------------------------------------------------------------------------- */
/* InvBiCGStab_a, lib/actions/ferm/invert/invbicgstab.cc:10-202                */
/* ------------------------------------------------------------------------- */
/* Returns n_count (MaxBiCGStab if not converged), -1 on breakdown (the reference aborts). */
int orc_invbicgstab(orc_op* op, const double* chi, double* psi, double Rsd, int MaxIter, int isign, double* resid) {
  const orc_geom* g = op->g;
  size_t V = (size_t)g->V, n = (size_t)g->Vh*12, off = n;
  cplx* r  = (cplx*)calloc(V*12, sizeof(cplx));
  cplx* r0 = (cplx*)calloc(V*12, sizeof(cplx));
  cplx* p  = (cplx*)calloc(V*12, sizeof(cplx));
  cplx* v  = (cplx*)calloc(V*12, sizeof(cplx));
  cplx* t  = (cplx*)calloc(V*12, sizeof(cplx));
  cplx* ps = (cplx*)psi; const cplx* ch = (const cplx*)chi;
  int n_count = MaxIter; int convP = 0;
  *resid = 0.0;
  double chi_sq = norm2_odd(g, chi);
  double rsd_sq = Rsd*Rsd*chi_sq;
  orc_op_apply(op, (double*)r0, psi, isign);
#pragma omp parallel for
  for (size_t i = 0; i < n; ++i) { r[off+i] = c_sub(ch[off+i], r0[off+i]); r0[off+i] = r[off+i]; }
  cplx rho, rho_prev = {1,0}, alpha = {1,0}, omega = {1,0};
  for (int k = 1; k <= MaxIter && !convP; ++k) {
    rho = inner_odd(g, (double*)r0, (double*)r);
    if (rho.re == 0 && rho.im == 0) { n_count = -1; break; }
    cplx beta = c_mul(c_div(rho, rho_prev), c_div(alpha, omega));
#pragma omp parallel for
    for (size_t i = 0; i < n; ++i) {
      cplx tmp = c_sub(p[off+i], c_mul(omega, v[off+i]));
      p[off+i] = c_add(r[off+i], c_mul(beta, tmp));
    }
    orc_op_apply(op, (double*)v, (double*)p, isign);
    cplx ctmp = inner_odd(g, (double*)r0, (double*)v);
    if (ctmp.re == 0 && ctmp.im == 0) { n_count = -1; break; }
    alpha = c_div(rho, ctmp);
    rho_prev = rho;
#pragma omp parallel for
    for (size_t i = 0; i < n; ++i) r[off+i] = c_sub(r[off+i], c_mul(alpha, v[off+i]));
    orc_op_apply(op, (double*)t, (double*)r, isign);
    double t_norm = norm2_odd(g, (double*)t);
    if (t_norm == 0) { n_count = -1; break; }
    omega = inner_odd(g, (double*)t, (double*)r);
    omega.re /= t_norm; omega.im /= t_norm;
#pragma omp parallel for
    for (size_t i = 0; i < n; ++i) {
      cplx tmp = c_add(ps[off+i], c_mul(omega, r[off+i]));
      ps[off+i] = c_add(tmp, c_mul(alpha, p[off+i]));
      r[off+i] = c_sub(r[off+i], c_mul(omega, t[off+i]));
    }
    double r_norm = norm2_odd(g, (double*)r);
    if (r_norm < rsd_sq) { convP = 1; *resid = sqrt(r_norm); n_count = k; }
  }
  free(r); free(r0); free(p); free(v); free(t);
  return n_count;
}

/* LinOpSysSolverBiCGStab::operator(), syssolver_linop_bicgstab.h:57-95 */
int orc_solve_bicgstab(orc_op* op, const double* chi, double* psi, double Rsd, int MaxIter, double out[2]) {
  const orc_geom* g = op->g;
  size_t V = (size_t)g->V, n = (size_t)g->Vh*SPINOR, off = n;
  double dummy;
  int n_count = orc_invbicgstab(op, chi, psi, Rsd, MaxIter, +1, &dummy);
  double* tmp = (double*)calloc(V*SPINOR, sizeof(double));
  orc_op_apply(op, tmp, psi, +1);
  double s = 0;
#pragma omp parallel for reduction(+:s)
  for (size_t i = 0; i < n; ++i) { double t = chi[off+i] - tmp[off+i]; s += t*t; }
  out[0] = sqrt(s); out[1] = out[0]/sqrt(norm2_odd(g, chi));
  free(tmp);
  return n_count;
}

/* ------------------------------------------------------------------------- */
/* Full-lattice source preparation / solution reconstruction                   */
/* (lib/actions/ferm/qprop/eoprec_fermact_qprop.cc:41-80) -- "next" row (f1)    */
/* ------------------------------------------------------------------------- */
/* chi'_o = chi_o - D_oe A_ee^-1 chi_e  (the caller's unprec operator carries the -1/2 on D:
 * evenOddLinOp = -1/2 D_eo, eoprec_clover_linop_w.cc:98-133) */
void orc_qprop_prepare(orc_op* op, double* chi_prime, const double* chi) {
  int Vh = op->g->Vh; size_t n = (size_t)Vh*SPINOR;
  if (op->sym) {
    /* SymEvenOddPrecActQprop::operator() step (i), seoprec_fermact_qprop.cc:45-62:
     * chi' = L^-1 M_diag^-1 chi:  chi'_e = A_ee^-1 chi_e;  chi'_o = A_oo^-1 chi_o - M_oe chi'_e with
     * M_oe = A_oo^-1 (-1/2 Dslash)  (lib/seoprec_linop.h:188-204).  Both halves of chi_prime are written. */
    orc_clover_apply(op->g, chi_prime, chi, op->invclov, 0);
    orc_clover_apply(op->g, chi_prime, chi, op->invclov, 1);
    orc_op_dslash(op, op->tmp1, chi_prime, +1, 1);
#pragma omp parallel for
    for (size_t i = 0; i < n; ++i) op->tmp1[n+i] *= -0.5;
    orc_clover_apply(op->g, op->tmp2, op->tmp1, op->invclov, 1);
#pragma omp parallel for
    for (size_t i = 0; i < n; ++i) chi_prime[n+i] -= op->tmp2[n+i];
    return;
  }
  orc_clover_apply(op->g, op->tmp1, chi, op->invclov, 0);
  orc_op_dslash(op, op->tmp2, op->tmp1, +1, 1);
#pragma omp parallel for
  for (size_t i = 0; i < n; ++i) chi_prime[n+i] = chi[n+i] + 0.5*op->tmp2[n+i];
}
/* psi_e = A_ee^-1 (chi_e - D_eo psi_o), with D_eo = -1/2 Dslash */
void orc_qprop_reconstruct(orc_op* op, double* psi, const double* chi) {
  int Vh = op->g->Vh; size_t n = (size_t)Vh*SPINOR;
  if (op->sym) {
    /* step (ii), seoprec_fermact_qprop.cc:72-89; here chi is the PREPARED source chi' (its even half is the trivial
     * solution): psi_e = chi'_e - M_eo psi_o with M_eo = A_ee^-1 (-1/2 Dslash)  (lib/seoprec_linop.h:170-186) */
    orc_op_dslash(op, op->tmp1, psi, +1, 0);
#pragma omp parallel for
    for (size_t i = 0; i < n; ++i) op->tmp1[i] *= -0.5;
    orc_clover_apply(op->g, op->tmp2, op->tmp1, op->invclov, 0);
#pragma omp parallel for
    for (size_t i = 0; i < n; ++i) psi[i] = chi[i] - op->tmp2[i];
    return;
  }
  orc_op_dslash(op, op->tmp1, psi, +1, 0);
#pragma omp parallel for
  for (size_t i = 0; i < n; ++i) op->tmp2[i] = chi[i] + 0.5*op->tmp1[i];
  orc_clover_apply(op->g, psi, op->tmp2, op->invclov, 0);
}
/* Unpreconditioned operator on the full lattice: (A - 1/2 D) psi
 * (lib/eoprec_linop.h unprecLinOp: chi_e = A_ee psi_e + D_eo psi_o etc.) */
void orc_unprec_apply(orc_op* op, double* chi, const double* psi, int isign) {
  size_t n = (size_t)op->g->V*SPINOR;
  orc_clover_apply(op->g, chi, psi, op->clov, 0);
  orc_clover_apply(op->g, chi, psi, op->clov, 1);
  orc_op_dslash(op, op->tmp1, psi, isign, 0);
  orc_op_dslash(op, op->tmp1, psi, isign, 1);
#pragma omp parallel for
  for (size_t i = 0; i < n; ++i) chi[i] -= 0.5*op->tmp1[i];
}

/* ------------------------------------------------------------------------- */
/* CPU baseline support (bench.py): optional hook that routes the hopping term */
/* through the reference's own Dslash<double> (oracle/_ref), and a fixed-count  */
/* CG loop timer.                                                              */
/* ------------------------------------------------------------------------- */
typedef void (*orc_dslash_hook_fn)(void* handle, double* res, double* psi, double* packed_u, int isign, int source_cb);
static orc_dslash_hook_fn g_hook = 0;
static void* g_hook_handle = 0;
void orc_set_dslash_hook(orc_dslash_hook_fn fn, void* handle) { g_hook = fn; g_hook_handle = handle; }

static void bench_dslash(orc_op* op, double* chi, const double* psi, int isign, int cb) {
  if (g_hook) g_hook(g_hook_handle, chi, (double*)psi, op->packed_u, isign, 1 - cb);
  else orc_op_dslash(op, chi, psi, isign, cb);
}
static void bench_apply(orc_op* op, double* chi, const double* psi, int isign) {
  int Vh = op->g->Vh;
  bench_dslash(op, op->tmp1, psi, isign, 0);
  orc_clover_apply(op->g, op->tmp2, op->tmp1, op->invclov, 0);
  bench_dslash(op, op->tmp1, op->tmp2, isign, 1);
  orc_clover_apply(op->g, chi, psi, op->clov, 1);
  double* c = chi + (size_t)Vh*SPINOR; const double* t = op->tmp1 + (size_t)Vh*SPINOR;
  size_t n = (size_t)Vh*SPINOR;
#pragma omp parallel for
  for (size_t i = 0; i < n; ++i) c[i] += -0.25*t[i];
}

/* n_warm untimed + n_timed timed iterations of the InvCG2_a loop body (invcg2.cc:158-220), convergence test off.
 * Also times n_timed applications of M alone.  out = {secs per CG iteration, secs per M apply}. */
void orc_cg_bench(orc_op* op, const double* chi, int n_warm, int n_timed, double out[2]) {
  const orc_geom* g = op->g;
  size_t V = (size_t)g->V, n = (size_t)g->Vh*SPINOR, off = n;
  double* mp = (double*)calloc(V*SPINOR, sizeof(double));
  double* mmp = (double*)calloc(V*SPINOR, sizeof(double));
  double* p = (double*)calloc(V*SPINOR, sizeof(double));
  double* r = (double*)calloc(V*SPINOR, sizeof(double));
  double* psi = (double*)calloc(V*SPINOR, sizeof(double));
  memcpy(r + off, chi + off, n*sizeof(double));
  memcpy(p + off, chi + off, n*sizeof(double));
  double cp = norm2_odd(g, r), t0 = 0.0;
  for (int k = 0; k < n_warm + n_timed; ++k) {
#ifdef _OPENMP
    if (k == n_warm) t0 = omp_get_wtime();
#endif
    double c = cp;
    bench_apply(op, mp, p, +1);
    double d = norm2_odd(g, mp);
    bench_apply(op, mmp, mp, -1);
    double a = c/d;
#pragma omp parallel for
    for (size_t i = 0; i < n; ++i) r[off+i] -= a*mmp[off+i];
    cp = norm2_odd(g, r);
#pragma omp parallel for
    for (size_t i = 0; i < n; ++i) psi[off+i] += a*p[off+i];
    double b = cp/c;
#pragma omp parallel for
    for (size_t i = 0; i < n; ++i) p[off+i] = r[off+i] + b*p[off+i];
  }
#ifdef _OPENMP
  out[0] = (omp_get_wtime() - t0)/n_timed;
  t0 = omp_get_wtime();
  for (int k = 0; k < n_timed; ++k) bench_apply(op, mp, p, +1);
  out[1] = (omp_get_wtime() - t0)/n_timed;
#else
  out[0] = out[1] = 0.0;
#endif
  free(mp); free(mmp); free(p); free(r); free(psi);
}
