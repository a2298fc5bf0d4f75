// clover_setup.cuh -- one-time construction of the clover term on the GPU (rows a7/a8 of SURVEY.md section 8):
//   field strength  F_mu,nu = 1/8 (Q - Q^dag), Q = sum of the four plaquette leaves   (lib/meas/glue/mesfield.cc:44-74)
//   makeClov        A = diag_mass + sum c_mu,nu sigma_mu,nu F_mu,nu, packed triangular  (clover_term_qdp_w.h:398-553)
//   ldagdlinv       per-site LDL^dagger factorisation + inverse of both 6x6 blocks      (clover_term_qdp_w.h:619-846)
// These replace the QDP-JIT kernels ptx_make_clov / ptx_ldagdlinv (clover_term_ptx_w.h:780,1102).  Arithmetic is always
// in double (also for the fp32 engine); they run once per gauge field.
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
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    Z s = make_double2(0, 0);
#pragma unroll
    for (int k = 0; k < 3; ++k) s = zadd(s, zmul(a[i * 3 + k], b[k * 3 + j]));
    r[i * 3 + j] = s;
  }
}
__device__ __forceinline__ void mm_adj(Z* r, const Z* a, const Z* b) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    Z s = make_double2(0, 0);
#pragma unroll
    for (int k = 0; k < 3; ++k) s = zadd(s, zmul(a[i * 3 + k], zconj(b[j * 3 + k])));
    r[i * 3 + j] = s;
  }
}
__device__ __forceinline__ void adj_mm(Z* r, const Z* a, const Z* b) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    Z s = make_double2(0, 0);
#pragma unroll
    for (int k = 0; k < 3; ++k) s = zadd(s, zmul(zconj(a[k * 3 + i]), b[k * 3 + j]));
    r[i * 3 + j] = s;
  }
}

// The link U_mu at local coordinates c (on a split lattice c[3] may be -1 or Lt and c[2] may be -1 or Lz, both at once
// for the corner links; everything else wraps), as Chroma's state->getLinks() holds it: boundary phase included,
// anisotropy NOT included.  Coordinates are at most one step outside the local lattice.
template <typename R>
__device__ __forceinline__ void fetch_link(Z U[9], const CloverSetupArgs<R>& a, int mu, int cx, int cy, int cz, int ct) {
  const Geom& g = a.g;
  const int Lx = 2 * g.Lxh;
  cx = cx < 0 ? cx + Lx : (cx >= Lx ? cx - Lx : cx);
  cy = cy < 0 ? cy + g.Ly : (cy >= g.Ly ? cy - g.Ly : cy);
  if (!g.zsplit) cz = cz < 0 ? cz + g.Lz : (cz >= g.Lz ? cz - g.Lz : cz);
  if (!g.tsplit) ct = ct < 0 ? ct + g.Lt : (ct >= g.Lt ? ct - g.Lt : ct);
  // parity of the site: ghost slices keep the parity of their global coordinate; local extents are even, so
  // (t = -1) and (t = Lt) have the parity of an odd / even t respectively (same for z).
  const int par = (cx + cy + cz + ct + 4) & 1;
  if (g.zsplit && (cz < 0 || cz >= g.Lz)) {
    const int face = cz < 0 ? 0 : 1;
    const size_t sze = (size_t)g.Lxh * g.Ly * (g.Lt + 2);
    const int f = ((ct + 1) * g.Ly + cy) * g.Lxh + cx / 2;
    const Cx<R>* p = a.ghost_links_z + ((size_t)((face * 4 + mu) * 2 + par) * 9) * sze + f;
#pragma unroll
    for (int k = 0; k < 9; ++k) { const Cx<R> v = p[(size_t)k * sze]; U[k] = make_double2((double)v.x, (double)v.y); }
    return;
  }
  if (g.tsplit && (ct < 0 || ct >= g.Lt)) {
    const int face = ct < 0 ? 0 : 1;
    const int s3 = (cz * g.Ly + cy) * g.Lxh + cx / 2;
    const Cx<R>* p = a.ghost_links + ((size_t)((face * 4 + mu) * 2 + par) * 9) * g.S3h + s3;
#pragma unroll
    for (int k = 0; k < 9; ++k) { const Cx<R> v = p[(size_t)k * g.S3h]; U[k] = make_double2((double)v.x, (double)v.y); }
    return;
  }
  const int idx = ((ct * g.Lz + cz) * g.Ly + cy) * g.Lxh + cx / 2;
  const int NG = a.recon12 ? 6 : 9;
  const Cx<R>* p = a.gauge + ((size_t)(mu * 2 + par) * NG) * g.Vh + idx;
#pragma unroll
  for (int k = 0; k < 9; ++k) if (k < NG) { const Cx<R> v = __ldg(p + (size_t)k * g.Vh); U[k] = make_double2((double)v.x, (double)v.y); }
  if (a.recon12) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int c1 = (c + 1) % 3, c2 = (c + 2) % 3;
      U[6 + c] = zconj(zsub(zmul(U[c1], U[3 + c2]), zmul(U[c2], U[3 + c1])));
    }
    if (mu == 3 && a.bc_t == -1 && a.t_is_last && ct == g.Lt - 1) {
#pragma unroll
      for (int k = 0; k < 9; ++k) { U[k].x = -U[k].x; U[k].y = -U[k].y; }
    }
  } else {
    const double s = a.inv_aniso[mu];
#pragma unroll
    for (int k = 0; k < 9; ++k) { U[k].x *= s; U[k].y *= s; }
  }
}

// Field strength of ONE plane (MU < NU) at one site, times the clover coefficient of the plane:
// F = coef/8 (Q - Q^dag), Q = sum of the four leaves (mesfield.cc:44-74; getCloverCoeff, clover_term_qdp_w.h:1524-1544).
// At most two links and two products are live at a time (the one-thread-per-site version kept seven 3x3 matrices and
// the six F planes in local memory).
template <typename R, int MU, int NU>
__device__ __forceinline__ void field_strength_plane(Z Fout[9], const CloverSetupArgs<R>& a, int x, int y, int z, int t) {
  constexpr int ex = (MU == 0), ey = (MU == 1), ez = (MU == 2), et = (MU == 3);     // unit vector mu
  constexpr int fx = (NU == 0), fy = (NU == 1), fz = (NU == 2), ft = (NU == 3);     // unit vector nu
#define LNK(U, dir, sm, sn) fetch_link<R>(U, a, dir, x + (sm) * ex + (sn) * fx, y + (sm) * ey + (sn) * fy, z + (sm) * ez + (sn) * fz, t + (sm) * et + (sn) * ft)
  Z A[9], B[9], t1[9], t2[9], Q[9];
  // leaf 1: U_mu(x) U_nu(x+mu) U_mu(x+nu)^dag U_nu(x)^dag
  LNK(A, MU, 0, 0); LNK(B, NU, 1, 0); mm(t1, A, B);
  LNK(A, NU, 0, 0); LNK(B, MU, 0, 1); mm(t2, A, B);            // t2 = U_nu(x) U_mu(x+nu)
  mm_adj(Q, t1, t2);
  // leaf 2: U_mu(x-mu)^dag U_nu(x-mu-nu)^dag U_mu(x-mu-nu) U_nu(x-nu)
  LNK(A, NU, -1, -1); LNK(B, MU, -1, 0); mm(t1, A, B);         // U_nu(x-mu-nu) U_mu(x-mu)
  LNK(A, MU, -1, -1); LNK(B, NU, 0, -1); mm(t2, A, B);         // U_mu(x-mu-nu) U_nu(x-nu)
  adj_mm(A, t1, t2);
#pragma unroll
  for (int k = 0; k < 9; ++k) Q[k] = zadd(Q[k], A[k]);
  // leaf 3: U_nu(x-nu)^dag U_mu(x-nu) U_nu(x-nu+mu) U_mu(x)^dag
  LNK(A, NU, 0, -1); LNK(B, MU, 0, -1); adj_mm(t1, A, B);
  LNK(A, NU, 1, -1); LNK(B, MU, 0, 0); mm_adj(t2, A, B);
  mm(A, t1, t2);
#pragma unroll
  for (int k = 0; k < 9; ++k) Q[k] = zadd(Q[k], A[k]);
  // leaf 4: U_nu(x) U_mu(x-mu+nu)^dag U_nu(x-mu)^dag U_mu(x-mu)
  LNK(A, NU, 0, 0); LNK(B, MU, -1, 1); mm_adj(t1, A, B);
  LNK(A, NU, -1, 0); LNK(B, MU, -1, 0); adj_mm(t2, A, B);
  mm(A, t1, t2);
#pragma unroll
  for (int k = 0; k < 9; ++k) Q[k] = zadd(Q[k], A[k]);
#undef LNK
  const double coef = (a.aniso && (MU == a.t_dir || NU == a.t_dir)) ? a.ct : a.cr;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const Z d = zsub(Q[i * 3 + j], zconj(Q[j * 3 + i]));
      Fout[i * 3 + j] = make_double2(0.125 * d.x * coef, 0.125 * d.y * coef);
    }
}

// One entry of the packed clover term of a site from its six field-strength planes (makeClovSiteLoop,
// clover_term_qdp_w.h:416-519).  q = 0..2: the diagonal pairs (2q, 2q+1) of chiral block b; q = 3..17: off-diagonal k = q-3.
// F(p, e) reads element e of plane p of this site.  Planes: 0 (0,1), 1 (0,2), 2 (0,3), 3 (1,2), 4 (1,3), 5 (2,3).
template <typename FS>
__device__ __forceinline__ Z clover_entry(int b, int q, double diag_mass, const FS& F) {
  if (q < 3) {
    double d[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i6 = 2 * q + h, i = i6 % 3;          // diag index 0..5 -> colour i, upper (i6 < 3) or lower half
      const Z f5 = F(5, i * 3 + i), f0 = F(0, i * 3 + i);
      const double im = (b == 0) ? (f5.y - f0.y) : (f5.y + f0.y);
      const bool plus = (b == 0) ? (i6 < 3) : (i6 >= 3);
      d[h] = diag_mass + (plus ? im : -im);
    }
    return make_double2(d[0], d[1]);
  }
  const int k = q - 3;
  // colour-colour entries inside the upper (k = 0,1,2) and lower (k = 9,13,14) 3x3 sub-blocks
  if (k < 3 || k == 9 || k == 13 || k == 14) {
    const int kk = k < 3 ? k : (k == 9 ? 0 : k - 12);
    const int i = kk == 0 ? 1 : 2, j = kk == 2 ? 1 : 0;
    const Z f0 = F(0, i * 3 + j), f5 = F(5, i * 3 + j);
    Z v = (b == 0) ? ztimesI(zsub(f0, f5)) : ztimesI(zadd(f5, f0));
    if (k >= 3) { v.x = -v.x; v.y = -v.y; }
    return v;
  }
  // entries coupling the lower to the upper half: eij = (i+3)(i+2)/2 + j
  const int i = k < 6 ? 0 : (k < 9 ? 1 : 2);
  const int j = k - (i == 0 ? 3 : (i == 1 ? 6 : 10));
  const Z Em = zadd(ztimesI(F(2, i * 3 + j)), F(4, i * 3 + j));
  const Z Bm = zsub(ztimesI(F(3, i * 3 + j)), F(1, i * 3 + j));
  return (b == 0) ? zsub(Bm, Em) : zadd(Em, Bm);
}

// CTA = CLOV_SITES sites x 6 planes: warp p computes plane p of the field strength for 32 consecutive sites (6x the
// parallelism of one thread per site, no local-memory arrays), the planes meet in shared memory, and the 192 threads
// then assemble the 36 output planes, 6 each.  48^3x96: 45 ms -> see profiles/ (per parity).
constexpr int CLOV_SITES = 32;
#ifndef B200_CLOV_MINB
#define B200_CLOV_MINB 2   // CTAs/SM the register allocator plans for.  B200, 48^3x96: 1 (230 registers, no spills) 56 ms per parity; 2 (168 registers, ~400 B of spills) 28 ms -- the kernel is latency-bound, occupancy wins
#endif
struct SmemF {
  const Z* f; int lane;
  __device__ __forceinline__ Z operator()(int plane, int e) const { return f[(plane * 9 + e) * CLOV_SITES + lane]; }
};
template <typename R>
__global__ void __launch_bounds__(CLOV_SITES * 6, B200_CLOV_MINB) make_clover_kernel(const CloverSetupArgs<R> a) {
  __shared__ Z Fs[6 * 9 * CLOV_SITES];
  const Geom& g = a.g;
  const int lane = threadIdx.x, plane = threadIdx.y;
  const int idx = blockIdx.x * CLOV_SITES + lane;
  const bool active = idx < g.Vh;
  if (active) {
    int q = idx;
    const int xh = q % g.Lxh; q /= g.Lxh;
    const int y = q % g.Ly; q /= g.Ly;
    const int z = q % g.Lz;
    const int t = q / g.Lz;
    const int x = 2 * xh + ((y + z + t + a.parity) & 1);
    Z F[9];
    switch (plane) {       // uniform per warp
      case 0: field_strength_plane<R, 0, 1>(F, a, x, y, z, t); break;
      case 1: field_strength_plane<R, 0, 2>(F, a, x, y, z, t); break;
      case 2: field_strength_plane<R, 0, 3>(F, a, x, y, z, t); break;
      case 3: field_strength_plane<R, 1, 2>(F, a, x, y, z, t); break;
      case 4: field_strength_plane<R, 1, 3>(F, a, x, y, z, t); break;
      default: field_strength_plane<R, 2, 3>(F, a, x, y, z, t); break;
    }
#pragma unroll
    for (int e = 0; e < 9; ++e) Fs[(plane * 9 + e) * CLOV_SITES + lane] = F[e];
  }
  __syncthreads();
  if (!active) return;
  const SmemF F{Fs, lane};
  const size_t Vh = g.Vh;
#pragma unroll
  for (int n = 0; n < 6; ++n) {
    const int p = plane * 6 + n;                  // output plane 0..35 = block*18 + q
    const Z v = clover_entry(p / 18, p % 18, a.diag_mass, F);
    a.clov_out[(size_t)p * Vh + idx] = mk<R>((R)v.x, (R)v.y);
  }
}

// In-place inverse of one checkerboard's clover planes and tr_log[idx] = sum_i log|d_i| over both chiral blocks: what
// invclov.choles(cb) leaves behind (clover_term_qdp_w.h:560-571; the reference's site loop, :619-818, factorises A = L D L^dag
// and then solves for A^-1 one column at a time by forward and backward substitution).
//
// Here: one thread per (site, chiral block) -- 2 Vh threads, twice the parallelism of a thread per site -- and the whole
// 6x6 Hermitian block lives in registers: every loop below is fully unrolled, so the triangular index i(i-1)/2+j is a
// compile-time constant and no array is ever addressed dynamically (the round-1 kernel kept inv_offd[15] and v[6] in
// local memory).  The inverse is formed as an explicit congruence instead of 6 triangular solves:
//     A = L D L^dag   (left-looking, d_j real)
//     W = L^-1        (unit lower triangular, by forward substitution on the 15 strictly-lower entries)
//     A^-1 = W^dag D^-1 W,   (A^-1)_ij = sum_{k >= i} conj(W_ki) W_kj / d_k   for i >= j, W_kk = 1
// which needs no division beyond the six 1/d_k.  Arithmetic in double for both engine precisions.
constexpr __host__ __device__ int tri_idx(int i, int j) { return i * (i - 1) / 2 + j; }   // i > j
template <typename R>
__global__ void __launch_bounds__(CLOV_BLOCK) ldagdlinv_kernel(Cx<R>* __restrict__ tri, double* __restrict__ tr_log, int Vh_) {
  const int tid = blockIdx.x * CLOV_BLOCK + threadIdx.x;
  const int idx = tid >> 1, block = tid & 1;           // the two blocks of a site sit in adjacent lanes
  const bool active = idx < Vh_;
  const size_t Vh = Vh_;
  double tl = 0.0;
  if (active) {
    Cx<R>* p = tri + (size_t)(18 * block) * Vh + idx;
    double d[6]; Z l[15];
#pragma unroll
    for (int k = 0; k < 3; ++k) { const Cx<R> v = p[(size_t)k * Vh]; d[2 * k] = (double)v.x; d[2 * k + 1] = (double)v.y; }
#pragma unroll
    for (int k = 0; k < 15; ++k) { const Cx<R> o = p[(size_t)(3 + k) * Vh]; l[k] = make_double2((double)o.x, (double)o.y); }
    // ---- A = L D L^dag, in place: l[tri_idx(i,j)] becomes L_ij, d[j] the pivots
    double rd[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
#pragma unroll
      for (int k = 0; k < j; ++k) {
        const Z ljk = l[tri_idx(j, k)];
        d[j] -= (ljk.x * ljk.x + ljk.y * ljk.y) * d[k];
      }
      rd[j] = 1.0 / d[j];
      tl += log(fabs(d[j]));
#pragma unroll
      for (int i = j + 1; i < 6; ++i) {
        Z s = l[tri_idx(i, j)];
#pragma unroll
        for (int k = 0; k < j; ++k) {            // s -= L_ik d_k conj(L_jk)
          const Z lik = l[tri_idx(i, k)], ljk = l[tri_idx(j, k)];
          const double tr = (lik.x * ljk.x + lik.y * ljk.y) * d[k], ti = (lik.y * ljk.x - lik.x * ljk.y) * d[k];
          s.x -= tr; s.y -= ti;
        }
        l[tri_idx(i, j)] = make_double2(s.x * rd[j], s.y * rd[j]);
      }
    }
    // ---- W = L^-1: W_ij = -L_ij - sum_{j<k<i} L_ik W_kj
    Z w[15];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
#pragma unroll
      for (int i = j + 1; i < 6; ++i) {
        Z s = make_double2(-l[tri_idx(i, j)].x, -l[tri_idx(i, j)].y);
#pragma unroll
        for (int k = j + 1; k < i; ++k) {
          const Z lik = l[tri_idx(i, k)], wkj = w[tri_idx(k, j)];
          s.x -= lik.x * wkj.x - lik.y * wkj.y; s.y -= lik.x * wkj.y + lik.y * wkj.x;
        }
        w[tri_idx(i, j)] = s;
      }
    }
    // ---- A^-1 = W^dag D^-1 W (lower triangle), written straight back to the planes
    double od[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      double s = rd[i];
#pragma unroll
      for (int k = i + 1; k < 6; ++k) { const Z wki = w[tri_idx(k, i)]; s += (wki.x * wki.x + wki.y * wki.y) * rd[k]; }
      od[i] = s;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) p[(size_t)k * Vh] = mk<R>((R)od[2 * k], (R)od[2 * k + 1]);
#pragma unroll
    for (int i = 1; i < 6; ++i) {
#pragma unroll
      for (int j = 0; j < i; ++j) {
        Z s = make_double2(w[tri_idx(i, j)].x * rd[i], w[tri_idx(i, j)].y * rd[i]);      // k = i term: conj(W_ii = 1) W_ij / d_i
#pragma unroll
        for (int k = i + 1; k < 6; ++k) {       // + conj(W_ki) W_kj / d_k
          const Z wki = w[tri_idx(k, i)], wkj = w[tri_idx(k, j)];
          s.x += (wki.x * wkj.x + wki.y * wkj.y) * rd[k]; s.y += (wki.x * wkj.y - wki.y * wkj.x) * rd[k];
        }
        p[(size_t)(3 + tri_idx(i, j)) * Vh] = mk<R>((R)s.x, (R)s.y);
      }
    }
  }
  // tr_log of the site = sum over its two blocks (adjacent lanes)
  tl += __shfl_xor_sync(0xffffffffu, tl, 1);
  if (active && block == 0) tr_log[idx] = tl;
}

__global__ void sum_double_kernel(const double* x, size_t n, ReduceBuf red, double* dst);

}  // namespace b200
