// clover_setup.cuh -- one-time construction of the clover term on the GPU (rows a7/a8 of SURVEY.md section 8):
//   field strength  F_mu,nu = 1/8 (Q - Q^dag), Q = sum of the four plaquette leaves   (lib/meas/glue/mesfield.cc:44-74)
//   makeClov        A = diag_mass + sum c_mu,nu sigma_mu,nu F_mu,nu, packed triangular  (clover_term_qdp_w.h:398-553)
//   ldagdlinv       per-site LDL^dagger factorisation + inverse of both 6x6 blocks      (clover_term_qdp_w.h:619-846)
// These replace the QDP-JIT kernels ptx_make_clov / ptx_ldagdlinv (clover_term_ptx_w.h:780,1102).  One thread
// per site, arithmetic always in double (also for the fp32 engine); they run once per gauge field.
#pragma once
#include "common.cuh"
#include "reduce.cuh"

namespace b200 {

constexpr int CLOV_BLOCK = 64;

template <typename R>
struct CloverSetupArgs {
  const Cx<R>* gauge;       // engine gauge planes [4][2][NG][Vh]
  const Cx<R>* ghost_links; // T-split only: [2 faces (t=-1, t=Lt)][4][2][9][S3h], original links incl. phases
  const Cx<R>* ghost_links_z; // Z-split only: [2 faces (z=-1, z=Lz)][4][2][9][(Lt+2)*Ly*Lxh], rows t = -1 .. Lt (corners included)
  Cx<R>* clov_out;          // [36][Vh] of this parity
  double inv_aniso[4];      // undo the anisotropy factor folded into uncompressed links
  int recon12, bc_t, t_is_last;
  double diag_mass, cr, ct;
  int aniso, t_dir, parity;
  Geom g;
};

typedef double2 Z;
__device__ __forceinline__ Z zmul(Z a, Z b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ Z zconj(Z a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ Z zadd(Z a, Z b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ Z zsub(Z a, Z b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ Z ztimesI(Z a) { return make_double2(-a.y, a.x); }
__device__ __forceinline__ Z zdiv(Z a, Z b) {
  const double d = b.x * b.x + b.y * b.y;
  return make_double2((a.x * b.x + a.y * b.y) / d, (a.y * b.x - a.x * b.y) / d);
}
// r = a b, r = a b^dag, r = a^dag b on 3x3 complex matrices
__device__ __forceinline__ void mm(Z* r, const Z* a, const Z* b) {
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
    Z s = make_double2(0, 0);
    for (int k = 0; k < 3; ++k) s = zadd(s, zmul(a[i * 3 + k], b[k * 3 + j]));
    r[i * 3 + j] = s;
  }
}
__device__ __forceinline__ void mm_adj(Z* r, const Z* a, const Z* b) {
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
    Z s = make_double2(0, 0);
    for (int k = 0; k < 3; ++k) s = zadd(s, zmul(a[i * 3 + k], zconj(b[j * 3 + k])));
    r[i * 3 + j] = s;
  }
}
__device__ __forceinline__ void adj_mm(Z* r, const Z* a, const Z* b) {
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
    Z s = make_double2(0, 0);
    for (int k = 0; k < 3; ++k) s = zadd(s, zmul(zconj(a[k * 3 + i]), b[k * 3 + j]));
    r[i * 3 + j] = s;
  }
}

// The link U_mu at local coordinates c (on a split lattice c[3] may be -1 or Lt and c[2] may be -1 or Lz, both at once
// for the corner links; everything else wraps), as Chroma's state->getLinks() holds it: boundary phase included,
// anisotropy NOT included.
template <typename R>
__device__ void fetch_link(Z U[9], const CloverSetupArgs<R>& a, int mu, int cx, int cy, int cz, int ct) {
  const Geom& g = a.g;
  const int Lx = 2 * g.Lxh;
  cx = (cx + Lx) % Lx; cy = (cy + g.Ly) % g.Ly;
  if (!g.zsplit) cz = (cz + g.Lz) % g.Lz;
  if (!g.tsplit) ct = (ct + g.Lt) % g.Lt;
  // parity of the site: ghost slices keep the parity of their global coordinate; local extents are even, so
  // (t = -1) and (t = Lt) have the parity of an odd / even t respectively (same for z).
  const int par = (cx + cy + cz + ct + 4) & 1;
  if (g.zsplit && (cz < 0 || cz >= g.Lz)) {
    const int face = cz < 0 ? 0 : 1;
    const size_t sze = (size_t)g.Lxh * g.Ly * (g.Lt + 2);
    const int f = ((ct + 1) * g.Ly + cy) * g.Lxh + cx / 2;
    const Cx<R>* p = a.ghost_links_z + ((size_t)((face * 4 + mu) * 2 + par) * 9) * sze + f;
    for (int k = 0; k < 9; ++k) { const Cx<R> v = p[(size_t)k * sze]; U[k] = make_double2((double)v.x, (double)v.y); }
    return;
  }
  if (g.tsplit && (ct < 0 || ct >= g.Lt)) {
    const int face = ct < 0 ? 0 : 1;
    const int s3 = (cz * g.Ly + cy) * g.Lxh + cx / 2;
    const Cx<R>* p = a.ghost_links + ((size_t)((face * 4 + mu) * 2 + par) * 9) * g.S3h + s3;
    for (int k = 0; k < 9; ++k) { const Cx<R> v = p[(size_t)k * g.S3h]; U[k] = make_double2((double)v.x, (double)v.y); }
    return;
  }
  const int idx = ((ct * g.Lz + cz) * g.Ly + cy) * g.Lxh + cx / 2;
  const int NG = a.recon12 ? 6 : 9;
  const Cx<R>* p = a.gauge + ((size_t)(mu * 2 + par) * NG) * g.Vh + idx;
  for (int k = 0; k < NG; ++k) { const Cx<R> v = p[(size_t)k * g.Vh]; U[k] = make_double2((double)v.x, (double)v.y); }
  if (a.recon12) {
    for (int c = 0; c < 3; ++c) {
      const int c1 = (c + 1) % 3, c2 = (c + 2) % 3;
      U[6 + c] = zconj(zsub(zmul(U[c1], U[3 + c2]), zmul(U[c2], U[3 + c1])));
    }
    if (mu == 3 && a.bc_t == -1 && a.t_is_last && ct == g.Lt - 1)
      for (int k = 0; k < 9; ++k) { U[k].x = -U[k].x; U[k].y = -U[k].y; }
  } else {
    const double s = a.inv_aniso[mu];
    for (int k = 0; k < 9; ++k) { U[k].x *= s; U[k].y *= s; }
  }
}

template <typename R>
__global__ void __launch_bounds__(CLOV_BLOCK) make_clover_kernel(const CloverSetupArgs<R> a) {
  const Geom& g = a.g;
  const int idx = blockIdx.x * CLOV_BLOCK + threadIdx.x;
  if (idx >= g.Vh) return;
  int q = idx;
  const int xh = q % g.Lxh; q /= g.Lxh;
  const int y = q % g.Ly; q /= g.Ly;
  const int z = q % g.Lz;
  const int t = q / g.Lz;
  const int x = 2 * xh + ((y + z + t + a.parity) & 1);
  const int c0[4] = {x, y, z, t};

  Z F[6][9];
  int plane = 0;
#pragma unroll 1
  for (int mu = 0; mu < 3; ++mu) {
#pragma unroll 1
    for (int nu = mu + 1; nu < 4; ++nu, ++plane) {
      int em[4] = {0, 0, 0, 0}, en[4] = {0, 0, 0, 0};
      em[mu] = 1; en[nu] = 1;
#define LNK(U, dir, sm, sn) fetch_link<R>(U, a, dir, c0[0] + (sm) * em[0] + (sn) * en[0], c0[1] + (sm) * em[1] + (sn) * en[1], \
                                          c0[2] + (sm) * em[2] + (sn) * en[2], c0[3] + (sm) * em[3] + (sn) * en[3])
      Z A[9], B[9], C[9], D[9], t1[9], t2[9], Q[9];
      // leaf 1: U_mu(x) U_nu(x+mu) U_mu(x+nu)^dag U_nu(x)^dag
      LNK(A, mu, 0, 0); LNK(B, nu, 1, 0); LNK(C, mu, 0, 1); LNK(D, nu, 0, 0);
      mm(t1, A, B); mm(t2, D, C);           // t2 = U_nu(x) U_mu(x+nu)
      mm_adj(Q, t1, t2);
      // leaf 2: U_mu(x-mu)^dag U_nu(x-mu-nu)^dag U_mu(x-mu-nu) U_nu(x-nu)
      LNK(A, mu, -1, 0); LNK(B, nu, -1, -1); LNK(C, mu, -1, -1); LNK(D, nu, 0, -1);
      mm(t1, B, A);                          // U_nu(x-mu-nu) U_mu(x-mu)
      mm(t2, C, D);                          // U_mu(x-mu-nu) U_nu(x-nu)
      adj_mm(A, t1, t2);
      for (int k = 0; k < 9; ++k) Q[k] = zadd(Q[k], A[k]);
      // leaf 3: U_nu(x-nu)^dag U_mu(x-nu) U_nu(x-nu+mu) U_mu(x)^dag
      LNK(A, nu, 0, -1); LNK(B, mu, 0, -1); LNK(C, nu, 1, -1); LNK(D, mu, 0, 0);
      adj_mm(t1, A, B); mm_adj(t2, C, D);
      mm(A, t1, t2);
      for (int k = 0; k < 9; ++k) Q[k] = zadd(Q[k], A[k]);
      // leaf 4: U_nu(x) U_mu(x-mu+nu)^dag U_nu(x-mu)^dag U_mu(x-mu)
      LNK(A, nu, 0, 0); LNK(B, mu, -1, 1); LNK(C, nu, -1, 0); LNK(D, mu, -1, 0);
      mm_adj(t1, A, B); adj_mm(t2, C, D);
      mm(A, t1, t2);
      for (int k = 0; k < 9; ++k) Q[k] = zadd(Q[k], A[k]);
#undef LNK
      // F = 1/8 (Q - Q^dag), times the clover coefficient of this plane (getCloverCoeff, :1524-1544)
      const double coef = (a.aniso && (mu == a.t_dir || nu == a.t_dir)) ? a.ct : a.cr;
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        const Z d = zsub(Q[i * 3 + j], zconj(Q[j * 3 + i]));
        F[plane][i * 3 + j] = make_double2(0.125 * d.x * coef, 0.125 * d.y * coef);
      }
    }
  }

  // makeClovSiteLoop, clover_term_qdp_w.h:416-519
  double diag[2][6]; Z offd[2][15];
  for (int b = 0; b < 2; ++b) for (int i = 0; i < 6; ++i) diag[b][i] = a.diag_mass;
  for (int i = 0; i < 3; ++i) {
    const Z d0 = zsub(F[5][i * 3 + i], F[0][i * 3 + i]);
    diag[0][i] += d0.y; diag[0][i + 3] -= d0.y;
    const Z d1 = zadd(F[5][i * 3 + i], F[0][i * 3 + i]);
    diag[1][i] -= d1.y; diag[1][i + 3] += d1.y;
  }
  for (int i = 1; i < 3; ++i)
    for (int j = 0; j < i; ++j) {
      const int eij = i * (i - 1) / 2 + j, etmp = (i + 3) * (i + 2) / 2 + j + 3;
      offd[0][eij] = ztimesI(zsub(F[0][i * 3 + j], F[5][i * 3 + j]));
      offd[0][etmp] = make_double2(-offd[0][eij].x, -offd[0][eij].y);
      offd[1][eij] = ztimesI(zadd(F[5][i * 3 + j], F[0][i * 3 + j]));
      offd[1][etmp] = make_double2(-offd[1][eij].x, -offd[1][eij].y);
    }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      const int eij = (i + 3) * (i + 2) / 2 + j;
      const Z Em = zadd(ztimesI(F[2][i * 3 + j]), F[4][i * 3 + j]);
      const Z Bm = zsub(ztimesI(F[3][i * 3 + j]), F[1][i * 3 + j]);
      offd[0][eij] = zsub(Bm, Em);
      offd[1][eij] = zadd(Em, Bm);
    }
  const size_t Vh = g.Vh;
  for (int b = 0; b < 2; ++b) {
    Cx<R>* o = a.clov_out + (size_t)(18 * b) * Vh + idx;
    for (int k = 0; k < 3; ++k) o[(size_t)k * Vh] = mk<R>((R)diag[b][2 * k], (R)diag[b][2 * k + 1]);
    for (int k = 0; k < 15; ++k) o[(size_t)(3 + k) * Vh] = mk<R>((R)offd[b][k].x, (R)offd[b][k].y);
  }
}

// In-place inverse of the cb-0 clover planes; tr_log[idx] = sum_i log|d_i| over both blocks.
template <typename R>
__global__ void __launch_bounds__(CLOV_BLOCK) ldagdlinv_kernel(Cx<R>* __restrict__ tri, double* __restrict__ tr_log, int Vh_) {
  const int idx = blockIdx.x * CLOV_BLOCK + threadIdx.x;
  if (idx >= Vh_) return;
  const size_t Vh = Vh_;
  const int N = 6;
  double tl = 0.0;
  for (int block = 0; block < 2; ++block) {
    Cx<R>* p = tri + (size_t)(18 * block) * Vh + idx;
    double inv_d[6], diag_g[6]; Z inv_offd[15], v[6];
    for (int k = 0; k < 3; ++k) { const Cx<R> d = p[(size_t)k * Vh]; inv_d[2 * k] = (double)d.x; inv_d[2 * k + 1] = (double)d.y; }
    for (int k = 0; k < 15; ++k) { const Cx<R> o = p[(size_t)(3 + k) * Vh]; inv_offd[k] = make_double2((double)o.x, (double)o.y); }
    for (int j = 0; j < N; ++j) {
      for (int i = 0; i < j; ++i) {
        const int eji = j * (j - 1) / 2 + i;
        v[i] = zmul(make_double2(inv_d[i], 0.0), zconj(inv_offd[eji]));
      }
      v[j] = make_double2(inv_d[j], 0.0);
      for (int k = 0; k < j; ++k) v[j] = zsub(v[j], zmul(inv_offd[j * (j - 1) / 2 + k], v[k]));
      inv_d[j] = v[j].x;
      for (int k = j + 1; k < N; ++k) {
        const int ekj = k * (k - 1) / 2 + j;
        for (int l = 0; l < j; ++l) inv_offd[ekj] = zsub(inv_offd[ekj], zmul(inv_offd[k * (k - 1) / 2 + l], v[l]));
        inv_offd[ekj] = zdiv(inv_offd[ekj], v[j]);
      }
    }
    for (int i = 0; i < N; ++i) { diag_g[i] = 1.0 / inv_d[i]; tl += log(fabs(inv_d[i])); }
    for (int k = 0; k < N; ++k) {
      for (int i = 0; i < k; ++i) v[i] = make_double2(0, 0);
      v[k] = make_double2(diag_g[k], 0.0);
      for (int i = k + 1; i < N; ++i) {
        v[i] = make_double2(0, 0);
        for (int j = k; j < i; ++j)
          v[i] = zsub(v[i], zmul(zmul(inv_offd[i * (i - 1) / 2 + j], make_double2(inv_d[j], 0.0)), v[j]));
        v[i].x *= diag_g[i]; v[i].y *= diag_g[i];
      }
      for (int i = N - 2; i >= k; --i)
        for (int j = i + 1; j < N; ++j) v[i] = zsub(v[i], zmul(zconj(inv_offd[j * (j - 1) / 2 + i]), v[j]));
      inv_d[k] = v[k].x;
      for (int i = k + 1; i < N; ++i) inv_offd[i * (i - 1) / 2 + k] = v[i];
    }
    for (int k = 0; k < 3; ++k) p[(size_t)k * Vh] = mk<R>((R)inv_d[2 * k], (R)inv_d[2 * k + 1]);
    for (int k = 0; k < 15; ++k) p[(size_t)(3 + k) * Vh] = mk<R>((R)inv_offd[k].x, (R)inv_offd[k].y);
  }
  tr_log[idx] = tl;
}

__global__ void sum_double_kernel(const double* x, size_t n, ReduceBuf red, double* dst);

}  // namespace b200
