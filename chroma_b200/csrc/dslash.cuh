// dslash.cuh -- the hot kernels: Wilson hopping term with fused clover / solver epilogues.
//
// One thread per target-checkerboard site.  Per site the thread gathers 8 neighbour spinors
// (12 x 128-bit loads each, coalesced across the warp because idx is the fastest index of every
// plane), spin-projects them to half spinors while loading, multiplies by the 3x3 link
// (9 x 128-bit loads, or 6 with 12-real reconstruction), reconstructs and accumulates -- the same
// grouping of operations as the reference site loop (cpp_dslash_scalar_64bit.cc:105-213 with the site
// ops of cpp_dslash_scalar_64bit_c.h and su3_mult / su3_adj_mult of cpp_dslash_matvec64bit_c.h) --
// and then runs one of the fused epilogues:
//
//   EPI_DSLASH   out = D in                                       (Dslash<REAL>::operator())
//   EPI_AINV     out = A^-1 D in                                  (pass 1 of CloverSchur4D, cpp_clover_scalar_64bit.cc:137-237)
//   EPI_M        out = A x - 1/4 D in                             (pass 2, :248-377; eoprec_clover_linop_w.cc:167-171)
//   EPI_M_NORM   EPI_M and |out|^2                    -> CG d     (invcg2.cc:162-165)
//   EPI_M_CG     r -= a (A x - 1/4 D in), |r|^2       -> CG cp    (invcg2.cc:170-182; M^dag M p is never stored)
//   EPI_M_DOTR0  EPI_M and <r0|out>                   -> BiCGStab alpha   (invbicgstab.cc:101-114)
//   EPI_M_DOTX   EPI_M and <out|x>, |out|^2           -> BiCGStab omega   (invbicgstab.cc:126-140)
//   EPI_M_CGREL  EPI_M_CG whose finaliser also takes the reliable-update decisions (reliable_cg.cc:113-121)
//
// The EPI_M* epilogues come in three operator modes (template parameter MODE, DslashArgs::mmode):
//   MODE_ASYM       m = A x - 1/4 D in            EvenOddPrecCloverLinOp, eoprec_clover_linop_w.cc:142-187
//   MODE_SYM_PLUS   m = x - 1/4 A^-1 (D in)       SymEvenOddPrecCloverLinOp PLUS,  seoprec_clover_linop_w.cc:160-166, 175
//   MODE_SYM_MINUS  m = x - 1/4 D in              ... MINUS (A_oo^-1 was applied to the source first), :168-175
//
// Neighbour indices come from coordinate arithmetic (no shift table in HBM; replaces ShiftTable,
// shift_table_scalar.h:10-147).  The arithmetic is HBM-bound: 1320 flop per (48+8G) reals moved.
#pragma once
#include "common.cuh"
#include "reduce.cuh"

namespace b200 {

enum Epilogue { EPI_DSLASH = 0, EPI_AINV, EPI_M, EPI_M_NORM, EPI_M_CG, EPI_M_DOTR0, EPI_M_DOTX, EPI_M_CGREL };
enum OperatorMode { MODE_ASYM = 0, MODE_SYM_PLUS = 1, MODE_SYM_MINUS = 2 };

template <typename R>
struct DslashArgs {
  typedef Cx<R> C;
  const C* in;      // source checkerboard field (parity 1-parity)
  C* out;           // result (target parity); unused by EPI_M_CG
  const C* gauge;   // C[4][2][NG][Vh]
  const C* clov;    // clover planes of the TARGET parity (A_oo or A_ee^-1): C[36][Vh]
  const C* x;       // EPI_M*: the field A acts on (target parity)
  C* r;             // EPI_M_CG: residual, updated in place
  const C* r0;      // EPI_M_DOTR0: shadow residual
  // T-split ghost faces (half spinors, C[6][S3h]); only read when g.tsplit
  const C* ghost_fwd;   // (1 -/+ g3) psi(x+t) projected by the +t neighbour rank, for sites at t = Lt-1
  const C* ghost_bwd;   // U_t^dag (1 +/- g3) psi(x-t) projected AND multiplied by the -t rank, for sites at t = 0
  // Z-split ghost faces (C[6][SZh], face index (t*Ly+y)*Lxh+xh); only read when g.zsplit
  const C* ghost_zfwd;  // (1 -/+ g2) psi(x+z) from the +z neighbour rank, for sites at z = Lz-1
  const C* ghost_zbwd;  // U_z^dag (1 +/- g2) psi(x-z) from the -z neighbour rank, for sites at z = 0
  double* scal;     // device scalars (ScalarSlot)
  int* status;      // device status (StatusSlot)
  ReduceBuf red;
  Geom g;
  int parity;       // target parity
  int isign;        // +1: D, -1: D^dagger
  SiteBox box[5];   // target sites of this launch: the union of nbox boxes (whole lattice / interior / boundary pieces)
  int nbox;
  int nsites;       // total number of target sites of this launch
  int iter;         // solver iteration this launch belongs to (for the stop flag)
  int check_stop;   // 1: return immediately if status[ST_STOP] != 0
  int run_if;       // != 0: status slot that must be non-zero for this launch to do anything (predicated launch)
  // multi-RHS launches (dslash_mrhs_kernel): in/out/x/r/r0 point at right-hand side 0 of `nrhs` consecutive fields
  int nrhs;
  size_t fstride;   // elements between consecutive right-hand sides of a batched field (12*Vh)
  size_t gstride;   // same for the T ghost faces (6*S3h)
  size_t gstride_z; // ... and the Z ghost faces (6*SZh)
  int zc_sites;     // traversal order of box 0: != 0 => sweep t inside chunks of this many sites per time slice ...
  int zc_dz, zc_dy, zc_ncy;   // ... a chunk = zc_dz z-planes x zc_dy y-rows (zc_ncy = Ly / zc_dy chunks side by side in y)
  int mmode;        // OperatorMode of the EPI_M* epilogues (selects the kernel instantiation; single-RHS kernels only)
  L2Policy pol;     // L2 residency descriptors, created once per engine and passed as kernel arguments: they then live in the
                    // constant bank / uniform registers.  (ncu source view, round 2: with createpolicy inside the kernel the
                    // reducing epilogues kept the descriptors in vector registers and paid one R2UR in front of EVERY hinted
                    // load/store -- 284 extra instructions per warp, +12 % of the instruction count.)
  double twist;     // EPI_M*: isign * twisted_m -- m += twist * i gamma_5 x (eoprec_clover_linop_w.cc:174-184); 0 = no twisted term
};

// The `local`-th target site of a launch (local < a.nsites).  ZC: the batched kernels sweep box 0 in z-chunks.
template <typename R, bool ZC>
__device__ __forceinline__ int launch_site(const DslashArgs<R>& a, int local) {
  const Geom& g = a.g;
  const int n0 = box_count(g, a.box[0]);
  if (local < n0) {
    if (ZC && a.zc_sites) {
      // Traversal order of a batch: the t+-1 neighbours of a site are re-read one time slice later, and one slice of
      // nrhs spinors (127 MB for 12 at 48^3, fp64) does not survive in the 126 MB L2.  So the launch sweeps t inside
      // (z, y) chunks small enough that three slices of a chunk stay L2-resident (memory layout unchanged: only WHICH
      // site a thread takes changes).  mrhs_site is the division-free twin of this decode.
      const int nt = a.box[0].nt, per = a.zc_sites * nt;
      const int zc = local / per, rem = local - zc * per;
      const int t = rem / a.zc_sites, w = rem - t * a.zc_sites;
      const int cz = zc / a.zc_ncy, cy = zc - cz * a.zc_ncy;
      const int rowc = a.zc_dy * g.Lxh;
      const int zz = w / rowc, r2 = w - zz * rowc;
      return (((a.box[0].t0 + t) * g.Lz + a.box[0].z0 + cz * a.zc_dz + zz) * g.Ly + cy * a.zc_dy) * g.Lxh + r2;
    }
    return box_site(g, a.box[0], local);
  }
  local -= n0;
#pragma unroll
  for (int k = 1; k < 5; ++k) {
    if (k < a.nbox) {
      const int n = box_count(g, a.box[k]);
      if (local < n) return box_site(g, a.box[k], local);
      local -= n;
    }
  }
  return 0;
}

// The `local`-th target site among boxes kbegin.. of a launch (the boundary pieces of a fused split-lattice launch).
template <typename R>
__device__ __forceinline__ int launch_site_from(const DslashArgs<R>& a, int local, int kbegin) {
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    if (k >= kbegin && k < a.nbox) {
      const int n = box_count(a.g, a.box[k]);
      if (local < n) return box_site(a.g, a.box[k], local);
      local -= n;
    }
  }
  return 0;
}

// Checkerboard coordinates of a target site next to its cb2 index: the batched kernels decode them ONCE per thread
// (mrhs_site) and hand them to the staging code and to dslash_site*, instead of re-deriving them from idx.
struct SiteCoord { int idx, xh, y, z, t; };
// Launch-invariant divisors of the kernels' site decode (FastDiv, common.cuh), filled on the host per launch and passed
// as a kernel argument (single-RHS and batched kernels alike).
struct MrhsDiv { FastDiv lxh, ly, lz, zc, per, row, nz0, ncy, rowc; };
template <typename R>
inline MrhsDiv make_mrhs_div(const DslashArgs<R>& a) {
  MrhsDiv d;
  d.lxh = make_fastdiv(a.g.Lxh); d.ly = make_fastdiv(a.g.Ly); d.lz = make_fastdiv(a.g.Lz);
  d.zc = make_fastdiv(a.zc_sites); d.per = make_fastdiv(a.zc_sites * a.box[0].nt);
  d.row = make_fastdiv(a.g.Lxh * a.g.Ly); d.nz0 = make_fastdiv(a.box[0].nz);
  d.ncy = make_fastdiv(a.zc_ncy); d.rowc = make_fastdiv(a.zc_dy * a.g.Lxh);
  return d;
}
// coordinates of the site with cb2 index idx
__device__ __forceinline__ SiteCoord site_coord(const Geom& g, const MrhsDiv& dv, int idx) {
  SiteCoord s;
  s.idx = idx;
  const int q = fast_div(idx, dv.lxh); s.xh = idx - q * g.Lxh;
  const int q2 = fast_div(q, dv.ly); s.y = q - q2 * g.Ly;
  s.t = fast_div(q2, dv.lz); s.z = q2 - s.t * g.Lz;
  return s;
}
// The `local`-th target site of a batched launch, with its coordinates: same traversal order as launch_site<R, true>
// (box 0 swept in z-chunks), without a single hardware integer division on the box-0 path.  (Round 2: the generic
// decode + the old slot-by-slot staging loop cost ~1000 of the ~3000 instructions a warp issued per 32 sites, all of
// them in front of its first load.)
template <typename R>
__device__ __forceinline__ SiteCoord mrhs_site(const DslashArgs<R>& a, const MrhsDiv& dv, int local) {
  const Geom& g = a.g;
  const SiteBox& b = a.box[0];
  const int row = g.Lxh * g.Ly;
  SiteCoord s;
  if (local < row * b.nz * b.nt) {
    if (a.zc_sites) {
      // chunked order: chunk (cz, cy) of zc_dz z-planes x zc_dy y-rows, all time slices of a chunk before the next chunk
      const int zc = fast_div(local, dv.per), rem = local - zc * (a.zc_sites * b.nt);
      const int tt = fast_div(rem, dv.zc), w = rem - tt * a.zc_sites;
      const int cz = fast_div(zc, dv.ncy), cy = zc - cz * a.zc_ncy;
      const int rowc = a.zc_dy * g.Lxh;
      const int zz = fast_div(w, dv.rowc), r2 = w - zz * rowc;
      const int yy = fast_div(r2, dv.lxh);
      s.xh = r2 - yy * g.Lxh;
      s.y = cy * a.zc_dy + yy;
      s.z = b.z0 + cz * a.zc_dz + zz;
      s.t = b.t0 + tt;
      s.idx = ((s.t * g.Lz + s.z) * g.Ly + s.y) * g.Lxh + s.xh;
    } else {
      const int q = fast_div(local, dv.row), w = local - q * row;      // local = (tt*nz + zz)*row + w
      const int tt = fast_div(q, dv.nz0), zz = q - tt * b.nz;
      s.t = b.t0 + tt; s.z = b.z0 + zz;
      s.y = fast_div(w, dv.lxh); s.xh = w - s.y * g.Lxh;
      s.idx = (s.t * g.Lz + s.z) * row + w;
    }
  } else s = site_coord(g, dv, launch_site<R, false>(a, local));       // the other boxes of a boundary launch
  return s;
}

// Batched-kernel knobs (tuned on B200, profiles/r01_tune_mrhs_prefetch.txt):
#ifndef B200_MRHS_PREFETCH
#define B200_MRHS_PREFETCH 1    // fp64: neighbour spinors software-pipelined through shared memory (dslash_site_pf); 0: direct loads
#endif
#ifndef B200_MRHS_PREFETCH_F
#define B200_MRHS_PREFETCH_F 0  // fp32: the extra shared-memory traffic costs more than the latency it hides
#endif
#ifndef B200_MRHS_L1
#define B200_MRHS_L1 0          // 1: batched kernels let the neighbour spinors allocate in L1 (x/y neighbours of a warp overlap)
#endif
#ifndef B200_MRHS_DEPTH
#define B200_MRHS_DEPTH 1       // neighbour spinors in flight per warp (= 6 KB fp64 shared-memory buffers per warp): 1 or 2
#endif
#ifndef B200_MRHS_CLOVER_LATE
#define B200_MRHS_CLOVER_LATE 0 // depth 2 only: 1 = the clover block is not staged up front but copied over the dead link slots mid-way
#endif
template <typename R> struct MrhsPrefetch {
  static constexpr bool on = sizeof(R) == 4 ? (B200_MRHS_PREFETCH_F != 0) : (B200_MRHS_PREFETCH != 0);
  static constexpr int depth = B200_MRHS_DEPTH;
  static constexpr bool clover_late = on && depth == 2 && (B200_MRHS_CLOVER_LATE != 0);
  static_assert(B200_MRHS_DEPTH == 1 || B200_MRHS_DEPTH == 2, "the first B200_MRHS_DEPTH fetches must be x hops");
};

// ---- spin projection while loading: (1 + sg*gamma_MU) psi, upper two components -----------------
// SM: the spinor was prefetched into shared memory (plane stride = `stride` = 32), see dslash_site_pf.
template <typename R, int MU, bool MR = false, bool SM = false>
__device__ __forceinline__ void load_project(Cx<R> h0[3], Cx<R> h1[3], const Cx<R>* __restrict__ p, int stride, R sg, uint64_t keep) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const Cx<R> a0 = SM ? p[(0 * 3 + c) * stride] : MR ? (B200_MRHS_L1 ? ld_keep(p + (0 * 3 + c) * (size_t)stride, keep) : ld_keep_nol1(p + (0 * 3 + c) * (size_t)stride, keep)) : ld_keep(p + (0 * 3 + c) * (size_t)stride, keep);
    const Cx<R> a1 = SM ? p[(1 * 3 + c) * stride] : MR ? (B200_MRHS_L1 ? ld_keep(p + (1 * 3 + c) * (size_t)stride, keep) : ld_keep_nol1(p + (1 * 3 + c) * (size_t)stride, keep)) : ld_keep(p + (1 * 3 + c) * (size_t)stride, keep);
    const Cx<R> a2 = SM ? p[(2 * 3 + c) * stride] : MR ? (B200_MRHS_L1 ? ld_keep(p + (2 * 3 + c) * (size_t)stride, keep) : ld_keep_nol1(p + (2 * 3 + c) * (size_t)stride, keep)) : ld_keep(p + (2 * 3 + c) * (size_t)stride, keep);
    const Cx<R> a3 = SM ? p[(3 * 3 + c) * stride] : MR ? (B200_MRHS_L1 ? ld_keep(p + (3 * 3 + c) * (size_t)stride, keep) : ld_keep_nol1(p + (3 * 3 + c) * (size_t)stride, keep)) : ld_keep(p + (3 * 3 + c) * (size_t)stride, keep);
    if (MU == 0) {          // h0 = a0 + sg*i*a3, h1 = a1 + sg*i*a2
      h0[c] = mk<R>(a0.x - sg * a3.y, a0.y + sg * a3.x);
      h1[c] = mk<R>(a1.x - sg * a2.y, a1.y + sg * a2.x);
    } else if (MU == 1) {   // h0 = a0 - sg*a3, h1 = a1 + sg*a2
      h0[c] = mk<R>(a0.x - sg * a3.x, a0.y - sg * a3.y);
      h1[c] = mk<R>(a1.x + sg * a2.x, a1.y + sg * a2.y);
    } else if (MU == 2) {   // h0 = a0 + sg*i*a2, h1 = a1 - sg*i*a3
      h0[c] = mk<R>(a0.x - sg * a2.y, a0.y + sg * a2.x);
      h1[c] = mk<R>(a1.x + sg * a3.y, a1.y - sg * a3.x);
    } else {                // h0 = a0 + sg*a2, h1 = a1 + sg*a3
      h0[c] = mk<R>(a0.x + sg * a2.x, a0.y + sg * a2.y);
      h1[c] = mk<R>(a1.x + sg * a3.x, a1.y + sg * a3.y);
    }
  }
}

// ---- reconstruct the lower two spin components and accumulate ------------------------------------
template <typename R, int MU>
__device__ __forceinline__ void recons_acc(Cx<R> acc[12], const Cx<R> r0[3], const Cx<R> r1[3], R sg) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    acc[c].x += r0[c].x;     acc[c].y += r0[c].y;
    acc[3 + c].x += r1[c].x; acc[3 + c].y += r1[c].y;
    if (MU == 0) {          // r2 = -sg*i*r1, r3 = -sg*i*r0
      acc[6 + c].x += sg * r1[c].y; acc[6 + c].y -= sg * r1[c].x;
      acc[9 + c].x += sg * r0[c].y; acc[9 + c].y -= sg * r0[c].x;
    } else if (MU == 1) {   // r2 = sg*r1, r3 = -sg*r0
      acc[6 + c].x += sg * r1[c].x; acc[6 + c].y += sg * r1[c].y;
      acc[9 + c].x -= sg * r0[c].x; acc[9 + c].y -= sg * r0[c].y;
    } else if (MU == 2) {   // r2 = -sg*i*r0, r3 = +sg*i*r1
      acc[6 + c].x += sg * r0[c].y; acc[6 + c].y -= sg * r0[c].x;
      acc[9 + c].x -= sg * r1[c].y; acc[9 + c].y += sg * r1[c].x;
    } else {                // r2 = sg*r0, r3 = sg*r1
      acc[6 + c].x += sg * r0[c].x; acc[6 + c].y += sg * r0[c].y;
      acc[9 + c].x += sg * r1[c].x; acc[9 + c].y += sg * r1[c].y;
    }
  }
}

// ---- link load (18 reals, or 12 + third-row reconstruction) --------------------------------------
template <typename R, bool RECON12, bool MR = false>
__device__ __forceinline__ void load_link(Cx<R> U[9], const Cx<R>* __restrict__ p, int stride, uint64_t strm) {
  constexpr int NG = RECON12 ? 6 : 9;
#pragma unroll
  for (int k = 0; k < NG; ++k) U[k] = MR ? p[k * stride] /* staged in shared memory */ : ld_stream(p + k * (size_t)stride, strm);
  if (RECON12) {
    // row2 = conj(row0 x row1): exact for SU(3); non-unit factors (anisotropy, -1 boundary phase)
    // are carried separately by the caller (see load_gauge in api.cu).
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int c1 = (c + 1) % 3, c2 = (c + 2) % 3;
      Cx<R> t = csub(cmul(U[c1], U[3 + c2]), cmul(U[c2], U[3 + c1]));
      U[6 + c] = mk<R>(t.x, -t.y);
    }
  }
}

template <typename R, bool ADJ>
__device__ __forceinline__ void su3_mul(Cx<R> r0[3], Cx<R> r1[3], const Cx<R> U[9], const Cx<R> h0[3], const Cx<R> h1[3]) {
#pragma unroll
  for (int row = 0; row < 3; ++row) {
    Cx<R> s0 = mk<R>(0, 0), s1 = mk<R>(0, 0);
#pragma unroll
    for (int col = 0; col < 3; ++col) {
      if (!ADJ) { cmac(s0, U[row * 3 + col], h0[col]); cmac(s1, U[row * 3 + col], h1[col]); }
      else      { cmac_conj(s0, U[col * 3 + row], h0[col]); cmac_conj(s1, U[col * 3 + row], h1[col]); }
    }
    r0[row] = s0; r1[row] = s1;
  }
}

// One hop: acc += recons( U(or U^dag) * project(psi_nbr) ).
template <typename R, int MU, bool ADJ, bool RECON12, bool MR = false>
__device__ __forceinline__ void hop(Cx<R> acc[12], const Cx<R>* __restrict__ psi_nbr, const Cx<R>* link,
                                    int stride, int lstride, R sg, R scale, const L2Policy& pol) {
  Cx<R> h0[3], h1[3], U[9], r0[3], r1[3];
  load_project<R, MU, MR>(h0, h1, psi_nbr, stride, sg, pol.keep);
  load_link<R, RECON12, MR>(U, link, lstride, pol.stream);
  if (RECON12) {
#pragma unroll
    for (int c = 0; c < 3; ++c) { h0[c].x *= scale; h0[c].y *= scale; h1[c].x *= scale; h1[c].y *= scale; }
  }
  su3_mul<R, ADJ>(r0, r1, U, h0, h1);
  recons_acc<R, MU>(acc, r0, r1, sg);
}

// Hops across a rank boundary (split direction MU).  The ghost faces are written by a PEER GPU during this very launch
// (fused split-lattice kernel, halo.cuh), so they are read through the coherent path, never ld.global.nc.  Forward: the half spinor was already projected by the +mu
// neighbour rank, only the link multiply is left.  Backward: U^dag (1 +/- g_mu) psi was computed by the -mu neighbour
// rank (it owns that link): just reconstruct.  gp points at this site's entry of the ghost face, fs = face stride.
template <typename R, int MU, bool RECON12, bool MR>
__device__ __forceinline__ void ghost_hop_fwd(Cx<R> acc[12], const Cx<R>* __restrict__ gp, int fs, const Cx<R>* link, int lstride,
                                              R sg, R scale, const L2Policy& pol) {
  Cx<R> h0[3], h1[3], U[9], r0[3], r1[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) { h0[c] = ld_stream_rw(gp + (size_t)c * fs, pol.stream); h1[c] = ld_stream_rw(gp + (size_t)(3 + c) * fs, pol.stream); }
  load_link<R, RECON12, MR>(U, link, lstride, pol.stream);
  if (RECON12) {
#pragma unroll
    for (int c = 0; c < 3; ++c) { h0[c].x *= scale; h0[c].y *= scale; h1[c].x *= scale; h1[c].y *= scale; }
  }
  su3_mul<R, false>(r0, r1, U, h0, h1);
  recons_acc<R, MU>(acc, r0, r1, sg);
}
template <typename R, int MU>
__device__ __forceinline__ void ghost_hop_bwd(Cx<R> acc[12], const Cx<R>* __restrict__ gp, int fs, R sg, const L2Policy& pol) {
  Cx<R> r0[3], r1[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) { r0[c] = ld_stream_rw(gp + (size_t)c * fs, pol.stream); r1[c] = ld_stream_rw(gp + (size_t)(3 + c) * fs, pol.stream); }
  recons_acc<R, MU>(acc, r0, r1, sg);
}

// Per-direction scale factors used only with RECON12 (anisotropy * boundary sign), see api.cu.
struct LinkScale {
  double aniso[4];    // aniso_coeff[mu]
  int bc_t;           // +1 / -1: phase on links U_t at global t = Lt_global-1
  int t_is_last;      // 1 if this rank holds the global last time slice
};

// The Wilson hopping term for one target site.  (xh,y,z,t) are its checkerboard coordinates.
// MR (multi-RHS kernels): the site's 8 links have been staged in shared memory, slot (2*mu + backward)*NG + k of lane
// `sml` (slot stride 32); otherwise they stream from global memory.
template <typename R, bool RECON12, bool MR = false>
__device__ __forceinline__ void dslash_site(Cx<R> acc[12], const DslashArgs<R>& a, const LinkScale& ls, int idx, const L2Policy& pol,
                                            const Cx<R>* sml = nullptr, const SiteCoord* sc = nullptr) {
  typedef Cx<R> C;
  const Geom& g = a.g;
  const int stride = g.Vh;
  int xh, y, z, t;
  if (sc) { xh = sc->xh; y = sc->y; z = sc->z; t = sc->t; }      // batched kernels: decoded once by mrhs_site
  else {
    int q = idx;
    xh = q % g.Lxh; q /= g.Lxh;
    y = q % g.Ly;   q /= g.Ly;
    z = q % g.Lz;
    t = q / g.Lz;
  }
  const int p = a.parity;
  const int r = (y + z + t + p) & 1;      // x = 2*xh + r
  constexpr int NG = RECON12 ? 6 : 9;
  const size_t gplane = (size_t)NG * stride;
  const C* __restrict__ Uf = a.gauge + (size_t)p * gplane + idx;         // forward links live on the target parity
  const C* __restrict__ Ub = a.gauge + (size_t)(1 - p) * gplane;          // backward links on the source parity
  const size_t gmu = 2 * gplane;
  const C* __restrict__ in = a.in;
  const R s = (R)a.isign;
  const int ls_ = MR ? 32 : stride;                 // link plane stride
  // link base pointers of the 8 hops
#define B200_LF(mu) (MR ? sml + (2 * (mu)) * NG * 32 : Uf + (mu) * gmu)
#define B200_LB(mu, nbr) (MR ? sml + (2 * (mu) + 1) * NG * 32 : Ub + (mu) * gmu + (nbr))

#pragma unroll
  for (int k = 0; k < 12; ++k) acc[k] = mk<R>(0, 0);

  // x direction: neighbours have the same idx or idx +/- 1 within the row
  {
    const int xf = r ? (xh + 1 == g.Lxh ? idx - (g.Lxh - 1) : idx + 1) : idx;
    const int xb = r ? idx : (xh == 0 ? idx + (g.Lxh - 1) : idx - 1);
    hop<R, 0, false, RECON12, MR>(acc, in + xf, B200_LF(0), stride, ls_, -s, (R)ls.aniso[0], pol);
    hop<R, 0, true, RECON12, MR>(acc, in + xb, B200_LB(0, xb), stride, ls_, s, (R)ls.aniso[0], pol);
  }
  {
    const int yf = (y + 1 == g.Ly) ? idx - (g.Ly - 1) * g.Lxh : idx + g.Lxh;
    const int yb = (y == 0) ? idx + (g.Ly - 1) * g.Lxh : idx - g.Lxh;
    hop<R, 1, false, RECON12, MR>(acc, in + yf, B200_LF(1), stride, ls_, -s, (R)ls.aniso[1], pol);
    hop<R, 1, true, RECON12, MR>(acc, in + yb, B200_LB(1, yb), stride, ls_, s, (R)ls.aniso[1], pol);
  }
  {
    const int sz = g.Ly * g.Lxh;
    const bool last = (z + 1 == g.Lz), first = (z == 0);
    const int zf = last ? idx - (g.Lz - 1) * sz : idx + sz;
    const int zb = first ? idx + (g.Lz - 1) * sz : idx - sz;
    const int fz = (t * g.Ly + y) * g.Lxh + xh;     // index of this site on a Z face
    if (g.zsplit && last) ghost_hop_fwd<R, 2, RECON12, MR>(acc, a.ghost_zfwd + fz, g.SZh, B200_LF(2), ls_, -s, (R)ls.aniso[2], pol);
    else hop<R, 2, false, RECON12, MR>(acc, in + zf, B200_LF(2), stride, ls_, -s, (R)ls.aniso[2], pol);
    if (g.zsplit && first) ghost_hop_bwd<R, 2>(acc, a.ghost_zbwd + fz, g.SZh, s, pol);
    else hop<R, 2, true, RECON12, MR>(acc, in + zb, B200_LB(2, zb), stride, ls_, s, (R)ls.aniso[2], pol);
  }
  {
    const int st = g.S3h;
    const bool last = (t + 1 == g.Lt), first = (t == 0);
    const int tf = last ? idx - (g.Lt - 1) * st : idx + st;
    const int tb = first ? idx + (g.Lt - 1) * st : idx - st;
    R scf = (R)ls.aniso[3], scb = (R)ls.aniso[3];
    if (RECON12) {   // the -1 of an antiperiodic boundary sits on U_t(t_global = last): forward hop from the last
                     // slice, backward hop into slice 0 (its link lives on the last slice)
      if (ls.t_is_last && last) scf *= (R)ls.bc_t;
      if (ls.t_is_last && first && !g.tsplit) scb *= (R)ls.bc_t;
    }
    if (g.tsplit && last) ghost_hop_fwd<R, 3, RECON12, MR>(acc, a.ghost_fwd + (idx - (g.Lt - 1) * st), st, B200_LF(3), ls_, -s, scf, pol);
    else hop<R, 3, false, RECON12, MR>(acc, in + tf, B200_LF(3), stride, ls_, -s, scf, pol);
    if (g.tsplit && first) ghost_hop_bwd<R, 3>(acc, a.ghost_bwd + idx, st, s, pol);
    else hop<R, 3, true, RECON12, MR>(acc, in + tb, B200_LB(3, tb), stride, ls_, s, scb, pol);
  }
#undef B200_LF
#undef B200_LB
}

// ---- batched kernels: the hopping term with the neighbour spinors software-pipelined through shared memory ----------
// The batched kernel is bound by latency x occupancy: a warp walks 8 dependent "load 12 complex numbers,
// wait, multiply" rounds, and the registers that would hold a second hop's loads in flight are the ones that limit the
// warps per SM.  So the NEXT hop's neighbour spinor is fetched with cp.async into a per-warp shared-memory buffer
// (12 planes x 32 lanes) while the current hop is multiplied: twice the bytes in flight per warp, and no register is
// tied up while they fly.  Every lane reads back only what it copied itself, so cp.async.wait_group is all the
// synchronisation needed.  `sp` = this lane's column of the warp's buffer; after the last hop the buffer receives
// `after` (the x operand of the EPI_M* epilogues) if given.  Same arithmetic, same order as dslash_site.
template <typename R>
__device__ __forceinline__ void prefetch_spinor(Cx<R>* sp, const Cx<R>* __restrict__ p, int stride, uint64_t pol) {
#pragma unroll
  for (int k = 0; k < 12; ++k) cp_async_hint<B200_MRHS_L1 != 0>(sp + k * 32, p + (size_t)k * stride, pol);
  cp_async_commit();
}
// cp.async bookkeeping of the pipelined hops: EVERY slot k = 0..8+depth-1 (8 hops, then the epilogue operand) commits
// exactly one group -- an empty one if there is nothing to fetch (ghost hop, no operand, past the end) -- so that
// "all but the depth-1 most recent groups have landed" always means "slot k has landed" at hop k.
template <int PENDING> __device__ __forceinline__ void cp_async_wait_pending() { asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory"); }
template <typename R, int MU, bool ADJ, bool RECON12>
__device__ __forceinline__ void hop_pf(Cx<R> acc[12], Cx<R>* sp, const Cx<R>* link, R sg, R scale, const L2Policy& pol,
                                       const Cx<R>* __restrict__ next, int stride, uint64_t next_pol) {
  Cx<R> h0[3], h1[3], U[9], r0[3], r1[3];
  cp_async_wait_pending<MrhsPrefetch<R>::depth - 1>();
  load_project<R, MU, true, true>(h0, h1, sp, 32, sg, 0);
  // the buffer is free again: start the fetch of the hop `depth` slots ahead into it
  if (next) prefetch_spinor<R>(sp, next, stride, next_pol); else cp_async_commit();
  load_link<R, RECON12, true>(U, link, 32, pol.stream);
  if (RECON12) {
#pragma unroll
    for (int c = 0; c < 3; ++c) { h0[c].x *= scale; h0[c].y *= scale; h1[c].x *= scale; h1[c].y *= scale; }
  }
  su3_mul<R, ADJ>(r0, r1, U, h0, h1);
  recons_acc<R, MU>(acc, r0, r1, sg);
}
// index of the +x / -x neighbour of a target site: the first hops of dslash_site_pf (the batched kernel issues their
// fetches early, next to the staging copies)
__device__ __forceinline__ int xfwd_neighbour(const Geom& g, const SiteCoord& c, int parity) {
  const int r = (c.y + c.z + c.t + parity) & 1;
  return r ? (c.xh + 1 == g.Lxh ? c.idx - (g.Lxh - 1) : c.idx + 1) : c.idx;
}
__device__ __forceinline__ int xbwd_neighbour(const Geom& g, const SiteCoord& c, int parity) {
  const int r = (c.y + c.z + c.t + parity) & 1;
  return r ? c.idx : (c.xh == 0 ? c.idx + (g.Lxh - 1) : c.idx - 1);
}
// the first `depth` fetches of a site (slots 0 and 1 are the x hops, never ghost hops); `sp` = the lane's column of
// the warp's first buffer, consecutive buffers 12*32 elements apart
template <typename R>
__device__ __forceinline__ void dslash_site_pf_begin(const Cx<R>* __restrict__ in, const Geom& g, const SiteCoord& c, int parity, Cx<R>* sp, uint64_t keep) {
  prefetch_spinor<R>(sp, in + xfwd_neighbour(g, c, parity), g.Vh, keep);
  if (MrhsPrefetch<R>::depth > 1) prefetch_spinor<R>(sp + 12 * 32, in + xbwd_neighbour(g, c, parity), g.Vh, keep);
}
// The hops of one site with the neighbour spinors pipelined through the warp's `depth` shared-memory buffers; the
// caller has run dslash_site_pf_begin.  Slot k lives in buffer k % depth; the fetch of slot k + depth goes into the
// buffer hop k has just read.
// H0, H1: only hops H0 <= k < H1 are done (the clover-late variant of the batched kernel splits a site's hops in two).
template <typename R, bool RECON12, int H0 = 0, int H1 = 8>
__device__ __forceinline__ void dslash_site_pf(Cx<R> acc[12], const DslashArgs<R>& a, const LinkScale& ls, const SiteCoord& c, const L2Policy& pol,
                                               const Cx<R>* sml, Cx<R>* sp, const Cx<R>* after) {
  typedef Cx<R> C;
  constexpr int D = MrhsPrefetch<R>::depth;
  const Geom& g = a.g;
  const int stride = g.Vh;
  const int idx = c.idx, xh = c.xh, y = c.y, z = c.z, t = c.t;
  constexpr int NG = RECON12 ? 6 : 9;
  const C* __restrict__ in = a.in;
  const R s = (R)a.isign;
#define B200_LF(mu) (sml + (2 * (mu)) * NG * 32)
#define B200_LB(mu) (sml + (2 * (mu) + 1) * NG * 32)
#define B200_BUF(k) (sp + ((k) % D) * (12 * 32))
  const int xb = xbwd_neighbour(g, c, a.parity);
  const int yf = (y + 1 == g.Ly) ? idx - (g.Ly - 1) * g.Lxh : idx + g.Lxh;
  const int yb = (y == 0) ? idx + (g.Ly - 1) * g.Lxh : idx - g.Lxh;
  const int sz = g.Ly * g.Lxh, st = g.S3h;
  const bool zl = (z + 1 == g.Lz), z0 = (z == 0), tl = (t + 1 == g.Lt), t0 = (t == 0);
  const int zf = zl ? idx - (g.Lz - 1) * sz : idx + sz;
  const int zb = z0 ? idx + (g.Lz - 1) * sz : idx - sz;
  const int tf = tl ? idx - (g.Lt - 1) * st : idx + st;
  const int tb = t0 ? idx + (g.Lt - 1) * st : idx - st;
  const bool gzf = g.zsplit && zl, gzb = g.zsplit && z0, gtf = g.tsplit && tl, gtb = g.tsplit && t0;   // hops served by ghost faces
  const int fz = (t * g.Ly + y) * g.Lxh + xh;
  R scf = (R)ls.aniso[3], scb = (R)ls.aniso[3];
  if (RECON12) {
    if (ls.t_is_last && tl) scf *= (R)ls.bc_t;
    if (ls.t_is_last && t0 && !g.tsplit) scb *= (R)ls.bc_t;
  }
  // what slot k fetches (nullptr: nothing -- ghost hop / no epilogue operand / past the end) and with which L2 policy
  const C* const nb[10] = {nullptr, in + xb, in + yf, in + yb, gzf ? nullptr : in + zf, gzb ? nullptr : in + zb,
                           gtf ? nullptr : in + tf, gtb ? nullptr : in + tb, after, nullptr};
#define B200_NEXT(k) nb[(k) + D], stride, ((k) + D == 8 ? pol.stream : pol.keep)
#define B200_HOP(k) if (H0 <= (k) && (k) < H1)
  if (H0 == 0) {
#pragma unroll
    for (int k = 0; k < 12; ++k) acc[k] = mk<R>(0, 0);
  }

  B200_HOP(0) hop_pf<R, 0, false, RECON12>(acc, B200_BUF(0), B200_LF(0), -s, (R)ls.aniso[0], pol, B200_NEXT(0));
  B200_HOP(1) hop_pf<R, 0, true, RECON12>(acc, B200_BUF(1), B200_LB(0), s, (R)ls.aniso[0], pol, B200_NEXT(1));
  B200_HOP(2) hop_pf<R, 1, false, RECON12>(acc, B200_BUF(2), B200_LF(1), -s, (R)ls.aniso[1], pol, B200_NEXT(2));
  B200_HOP(3) hop_pf<R, 1, true, RECON12>(acc, B200_BUF(3), B200_LB(1), s, (R)ls.aniso[1], pol, B200_NEXT(3));
  // a hop across a rank boundary reads its (already projected) half spinor from the ghost face: nothing was fetched
  // for its slot, so its buffer is free and the fetch `depth` slots ahead starts first
  B200_HOP(4) if (gzf) {
    if (nb[4 + D]) prefetch_spinor<R>(B200_BUF(4), B200_NEXT(4)); else cp_async_commit();
    ghost_hop_fwd<R, 2, RECON12, true>(acc, a.ghost_zfwd + fz, g.SZh, B200_LF(2), 32, -s, (R)ls.aniso[2], pol);
  } else hop_pf<R, 2, false, RECON12>(acc, B200_BUF(4), B200_LF(2), -s, (R)ls.aniso[2], pol, B200_NEXT(4));
  B200_HOP(5) if (gzb) {
    if (nb[5 + D]) prefetch_spinor<R>(B200_BUF(5), B200_NEXT(5)); else cp_async_commit();
    ghost_hop_bwd<R, 2>(acc, a.ghost_zbwd + fz, g.SZh, s, pol);
  } else hop_pf<R, 2, true, RECON12>(acc, B200_BUF(5), B200_LB(2), s, (R)ls.aniso[2], pol, B200_NEXT(5));
  B200_HOP(6) if (gtf) {
    if (nb[6 + D]) prefetch_spinor<R>(B200_BUF(6), B200_NEXT(6)); else cp_async_commit();
    ghost_hop_fwd<R, 3, RECON12, true>(acc, a.ghost_fwd + (idx - (g.Lt - 1) * st), st, B200_LF(3), 32, -s, scf, pol);
  } else hop_pf<R, 3, false, RECON12>(acc, B200_BUF(6), B200_LF(3), -s, scf, pol, B200_NEXT(6));
  B200_HOP(7) if (gtb) {
    if (nb[7 + D]) prefetch_spinor<R>(B200_BUF(7), B200_NEXT(7)); else cp_async_commit();
    ghost_hop_bwd<R, 3>(acc, a.ghost_bwd + idx, st, s, pol);
  } else hop_pf<R, 3, true, RECON12>(acc, B200_BUF(7), B200_LB(3), s, scb, pol, B200_NEXT(7));
#undef B200_LF
#undef B200_LB
#undef B200_BUF
#undef B200_NEXT
#undef B200_HOP
}

// Index (on the source checkerboard) of the backward neighbour in direction mu of a target site -- where the backward
// link lives.  Same arithmetic as dslash_site; used by the multi-RHS kernels to stage the links.
// `mu` is warp-uniform in the staging code, so the switch costs one uniform branch.
__device__ __forceinline__ int backward_neighbour(const Geom& g, const SiteCoord& c, int parity, int mu) {
  const int idx = c.idx;
  if (mu == 0) {
    const int r = (c.y + c.z + c.t + parity) & 1;
    return r ? idx : (c.xh == 0 ? idx + (g.Lxh - 1) : idx - 1);
  }
  if (mu == 1) return (c.y == 0) ? idx + (g.Ly - 1) * g.Lxh : idx - g.Lxh;
  const int sz = g.Ly * g.Lxh;
  if (mu == 2) return (c.z == 0) ? idx + (g.Lz - 1) * sz : idx - sz;
  return (c.t == 0) ? idx + (g.Lt - 1) * g.S3h : idx - g.S3h;
}

// ---- clover: one 6x6 Hermitian block times 6 complex --------------------------------------------
// out[i] = d_i in[i] + sum_{j<i} o_{k(i,j)} in[j] + sum_{j>i} conj(o_{k(j,i)}) in[j], k = i(i-1)/2+j
// (applySiteLoop, clover_term_qdp_w.h:1606-1634).  cl points at plane 0 of the block for this site.
template <typename R, bool MR = false>
__device__ __forceinline__ void clover_block(Cx<R> out[6], const Cx<R> in[6], const Cx<R>* cl, int stride, uint64_t strm) {
  const Cx<R> d01 = MR ? cl[0] : ld_stream(cl, strm);
  const Cx<R> d23 = MR ? cl[stride] : ld_stream(cl + (size_t)stride, strm);
  const Cx<R> d45 = MR ? cl[2 * stride] : ld_stream(cl + 2 * (size_t)stride, strm);
  out[0] = mk<R>(d01.x * in[0].x, d01.x * in[0].y);
  out[1] = mk<R>(d01.y * in[1].x, d01.y * in[1].y);
  out[2] = mk<R>(d23.x * in[2].x, d23.x * in[2].y);
  out[3] = mk<R>(d23.y * in[3].x, d23.y * in[3].y);
  out[4] = mk<R>(d45.x * in[4].x, d45.x * in[4].y);
  out[5] = mk<R>(d45.y * in[5].x, d45.y * in[5].y);
  int k = 0;
#pragma unroll
  for (int i = 1; i < 6; ++i) {
#pragma unroll
    for (int j = 0; j < i; ++j) {
      const Cx<R> o = MR ? cl[(3 + k) * stride] : ld_stream(cl + (size_t)(3 + k) * stride, strm);
      cmac(out[i], o, in[j]);
      cmac_conj(out[j], o, in[i]);
      ++k;
    }
  }
}

// ---- finalisers: run in one thread after the grid-wide (and cross-GPU) sum -----------------------
struct FinNone { __device__ void operator()(const double*) const {} };
// CG: d = |M p|^2  ->  a = c/d   (invcg2.cc:165,174)
struct FinCgD {
  double* scal;
  __device__ void operator()(const double* t) const { scal[S_D] = t[0]; scal[S_A] = scal[S_C] / t[0]; }
};
// CG: cp = |r|^2  ->  b = cp/c, c <- cp, convergence flag (invcg2.cc:182-216)
struct FinCgCp {
  double* scal; int* status; int iter; int check;
  __device__ void operator()(const double* t) const {
    const double cp = t[0], c = scal[S_C];
    scal[S_CP] = cp; scal[S_B] = cp / c; scal[S_C] = cp;
    if (check && status[ST_STOP] == 0 && cp <= scal[S_RSDSQ]) status[ST_STOP] = iter;
  }
};
// Reliable-update CG: cp = |r|^2 of the fp32 recurrence; decide on the device whether this iteration replaces the
// residual (updateR) and folds the partial solution into psi (updateX) -- reliable_cg.cc:113-121.  When it does,
// b, c and the convergence test are left to the finaliser of the replacement kernel (FinRelReplace, mixed.cuh).
struct FinRelCp {
  double* scal; int* status; int iter; int check;
  __device__ void operator()(const double* t) const {
    const double cp = t[0], rnorm = sqrt(cp);
    double maxrx = scal[S_MAXRX], maxrr = scal[S_MAXRR];
    const double r0 = scal[S_R0NORM], delta = scal[S_DELTA];
    if (rnorm > maxrx) maxrx = rnorm;
    if (rnorm > maxrr) maxrr = rnorm;
    scal[S_MAXRX] = maxrx; scal[S_MAXRR] = maxrr;
    const bool upd_x = (rnorm < delta * r0) && (r0 <= maxrx);
    const bool upd_r = ((rnorm < delta * maxrr) && (r0 <= maxrr)) || upd_x;
    status[ST_UPD_R] = upd_r ? 1 : 0; status[ST_UPD_X] = upd_x ? 1 : 0;
    scal[S_CP] = cp;
    if (!upd_r) {
      const double c = scal[S_C];
      scal[S_B] = cp / c; scal[S_C] = cp;
      if (check && status[ST_STOP] == 0 && cp < scal[S_RSDSQ]) status[ST_STOP] = iter;   // strict <, reliable_cg.cc:163
    }
  }
};
// BiCGStab: ctmp = <r0|v>  ->  alpha = rho/ctmp (invbicgstab.cc:106-114)
struct FinBiAlpha {
  double* scal; int* status;
  __device__ void operator()(const double* t) const {
    const double cr = t[0], ci = t[1];
    if (cr == 0.0 && ci == 0.0) { if (status[ST_BREAKDOWN] == 0) status[ST_BREAKDOWN] = 2; return; }
    const double rr = scal[S_RHO_RE], ri = scal[S_RHO_IM], d = cr * cr + ci * ci;
    scal[S_ALPHA_RE] = (rr * cr + ri * ci) / d;
    scal[S_ALPHA_IM] = (ri * cr - rr * ci) / d;
  }
};
// BiCGStab: omega = <t|r> / |t|^2 (invbicgstab.cc:130-140)
struct FinBiOmega {
  double* scal; int* status;
  __device__ void operator()(const double* t) const {
    const double tn = t[2];
    if (tn == 0.0) { if (status[ST_BREAKDOWN] == 0) status[ST_BREAKDOWN] = 3; return; }
    scal[S_OMEGA_RE] = t[0] / tn; scal[S_OMEGA_IM] = t[1] / tn;
  }
};

#ifndef B200_DSLASH_MINBLOCKS
#define B200_DSLASH_MINBLOCKS 1
#endif
#ifndef B200_DSLASH_MINBLOCKS_F
#define B200_DSLASH_MINBLOCKS_F 4   // fp32: 64-bit loads need 4 CTAs/SM in flight (tuned on B200: 1 -> 80 %, 3 -> 96 %, 4 -> 100 % of HBM peak)
#endif
// ---- fused epilogues for one target site (shared by the single- and multi-RHS kernels) ------------------------
// spx != nullptr (batched kernels with prefetch): the x operand of the EPI_M* epilogues was copied to shared memory
// (this lane's column, plane stride 32) behind the last hop.
template <typename R, int EPI, bool MR, int MODE = MODE_ASYM, bool XS = false>
__device__ __forceinline__ void site_epilogue(Cx<R> acc[12], const DslashArgs<R>& a, int idx, int stride, const L2Policy& pol, double red[3],
                                              const Cx<R>* smc = nullptr, const Cx<R>* spx = nullptr) {
  typedef Cx<R> C;
  // clover block b of this site: staged in shared memory (multi-RHS) or streamed from global memory
  const C* const cl0 = MR ? smc : a.clov + idx;
  const int cs = MR ? 32 : stride;
  if (EPI == EPI_DSLASH) {
#pragma unroll
    for (int k = 0; k < 12; ++k) st_stream(a.out + (size_t)k * stride + idx, acc[k], pol.stream);
  } else if (EPI == EPI_AINV) {
    C o[12];
#pragma unroll
    for (int b = 0; b < 2; ++b) clover_block<R, MR>(o + 6 * b, acc + 6 * b, cl0 + (size_t)(18 * b) * cs, cs, pol.stream);
#pragma unroll
    for (int k = 0; k < 12; ++k) st_stream(a.out + (size_t)k * stride + idx, o[k], pol.stream);
  } else {
    // all EPI_M* variants: m = A x - 1/4 D in.  Every load is issued before the first store (the stores are
    // volatile asm with a memory clobber, i.e. compiler barriers): a load placed after a store cannot be hoisted
    // and would cost one exposed DRAM round trip each.
    // Batched kernels (MR, 168 registers at most): the extra operand (r / r0) is loaded and consumed block by block --
    // all 12 of its components held across both clover blocks spilled 100 B per thread in EPI_M_CG.  Same sums, same order.
    C m[12], ex[MR ? 1 : 12];
    const R tw = (R)a.twist;
    const R cg_a = (EPI == EPI_M_CG || EPI == EPI_M_CGREL) ? (R)a.scal[S_A] : (R)0;
    if (XS) cp_async_wait_all();
    if (!MR && (EPI == EPI_M_CG || EPI == EPI_M_CGREL)) {
#pragma unroll
      for (int k = 0; k < 12; ++k) ex[k] = ld_stream_rw(a.r + (size_t)k * stride + idx, pol.stream);
    }
    if (!MR && EPI == EPI_M_DOTR0) {
#pragma unroll
      for (int k = 0; k < 12; ++k) ex[k] = ld_stream(a.r0 + (size_t)k * stride + idx, pol.stream);
    }
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      C xi[6], o[6], exb[6];
      if (MR && EPI == EPI_M_CG) {
#pragma unroll
        for (int k = 0; k < 6; ++k) exb[k] = ld_stream_rw(a.r + (size_t)(6 * b + k) * stride + idx, pol.stream);
      }
      if (MR && EPI == EPI_M_DOTR0) {
#pragma unroll
        for (int k = 0; k < 6; ++k) exb[k] = ld_stream(a.r0 + (size_t)(6 * b + k) * stride + idx, pol.stream);
      }
#pragma unroll
      for (int k = 0; k < 6; ++k) xi[k] = XS ? spx[(6 * b + k) * 32] : ld_stream(a.x + (size_t)(6 * b + k) * stride + idx, pol.stream);
      if (MODE == MODE_ASYM) clover_block<R, MR>(o, xi, cl0 + (size_t)(18 * b) * cs, cs, pol.stream);
      if (MODE == MODE_SYM_PLUS) clover_block<R, MR>(o, acc + 6 * b, cl0 + (size_t)(18 * b) * cs, cs, pol.stream);
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        if (MODE == MODE_ASYM) m[6 * b + k] = mk<R>(o[k].x - (R)0.25 * acc[6 * b + k].x, o[k].y - (R)0.25 * acc[6 * b + k].y);
        if (MODE == MODE_SYM_PLUS) m[6 * b + k] = mk<R>(xi[k].x - (R)0.25 * o[k].x, xi[k].y - (R)0.25 * o[k].y);
        if (MODE == MODE_SYM_MINUS) m[6 * b + k] = mk<R>(xi[k].x - (R)0.25 * acc[6 * b + k].x, xi[k].y - (R)0.25 * acc[6 * b + k].y);
        if (tw != (R)0) {   // twisted-mass term: i gamma_5 = +i on the upper chiral block (b = 0), -i on the lower one
          const R t = b == 0 ? tw : -tw;
          m[6 * b + k].x -= t * xi[k].y; m[6 * b + k].y += t * xi[k].x;
        }
        if (EPI == EPI_M_DOTX) {                              // <m|x>, |m|^2
          const C mm = m[6 * b + k];
          red[0] += (double)mm.x * xi[k].x + (double)mm.y * xi[k].y;
          red[1] += (double)mm.x * xi[k].y - (double)mm.y * xi[k].x;
          red[2] += (double)mm.x * mm.x + (double)mm.y * mm.y;
        }
        if (MR && EPI == EPI_M_CG) {                          // r -= a m, |r|^2 (batched: block by block)
          C rv = exb[k];
          rv.x -= cg_a * m[6 * b + k].x; rv.y -= cg_a * m[6 * b + k].y;
          red[0] += (double)rv.x * rv.x + (double)rv.y * rv.y;
          m[6 * b + k] = rv;
        }
        if (MR && EPI == EPI_M_NORM) red[0] += (double)m[6 * b + k].x * m[6 * b + k].x + (double)m[6 * b + k].y * m[6 * b + k].y;
        if (MR && EPI == EPI_M_DOTR0) {
          red[0] += (double)exb[k].x * m[6 * b + k].x + (double)exb[k].y * m[6 * b + k].y;
          red[1] += (double)exb[k].x * m[6 * b + k].y - (double)exb[k].y * m[6 * b + k].x;
        }
      }
    }
    if (EPI == EPI_M_CG || EPI == EPI_M_CGREL) {
      if (!MR) {
#pragma unroll
        for (int k = 0; k < 12; ++k) {
          C rv = ex[k];
          rv.x -= cg_a * m[k].x; rv.y -= cg_a * m[k].y;
          red[0] += (double)rv.x * rv.x + (double)rv.y * rv.y;
          m[k] = rv;
        }
      }
#pragma unroll
      for (int k = 0; k < 12; ++k) st_stream(a.r + (size_t)k * stride + idx, m[k], pol.stream);
    } else {
      if (!MR) {
#pragma unroll
        for (int k = 0; k < 12; ++k) {
          if (EPI == EPI_M_NORM) red[0] += (double)m[k].x * m[k].x + (double)m[k].y * m[k].y;
          if (EPI == EPI_M_DOTR0) {                             // <r0|m> = conj(r0) m
            red[0] += (double)ex[k].x * m[k].x + (double)ex[k].y * m[k].y;
            red[1] += (double)ex[k].x * m[k].y - (double)ex[k].y * m[k].x;
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 12; ++k) st_stream(a.out + (size_t)k * stride + idx, m[k], pol.stream);
    }
  }
}

template <typename R, int EPI, bool RECON12, int BLOCK, int MODE = MODE_ASYM>
__global__ void __launch_bounds__(BLOCK, (sizeof(R) == 4 ? B200_DSLASH_MINBLOCKS_F : B200_DSLASH_MINBLOCKS)) dslash_kernel(const DslashArgs<R> a, const LinkScale ls, const MrhsDiv dv) {
  typedef Cx<R> C;
  if (a.check_stop && (a.status[ST_STOP] != 0 || a.status[ST_BREAKDOWN] != 0)) return;
  if (a.run_if && a.status[a.run_if] == 0) return;
  const int stride = a.g.Vh;
  const int local = blockIdx.x * BLOCK + threadIdx.x;
  const bool active = local < a.nsites;
  double red[3] = {0.0, 0.0, 0.0};

  if (active) {
    // ZC: on lattices whose three live time slices of neighbour spinors outgrow the L2 budget (64^3: 75 MB) the launch
    // sweeps t inside z-chunks, like the batched kernels (zc_sites = 0 keeps the natural order)
    const SiteCoord sc = mrhs_site<R>(a, dv, local);      // division-free decode, same order as launch_site<R, true>
    const int idx = sc.idx;
    const L2Policy pol = a.pol;
    C acc[12];
    dslash_site<R, RECON12, false>(acc, a, ls, idx, pol, nullptr, &sc);
    site_epilogue<R, EPI, false, MODE>(acc, a, idx, stride, pol, red);
  }

  if (EPI == EPI_M_NORM) grid_reduce<1, BLOCK>(red, a.red, FinCgD{a.scal});
  if (EPI == EPI_M_CG) grid_reduce<1, BLOCK>(red, a.red, FinCgCp{a.scal, a.status, a.iter, a.check_stop});
  if (EPI == EPI_M_CGREL) grid_reduce<1, BLOCK>(red, a.red, FinRelCp{a.scal, a.status, a.iter, a.check_stop});
  if (EPI == EPI_M_DOTR0) grid_reduce<2, BLOCK>(red, a.red, FinBiAlpha{a.scal, a.status});
  if (EPI == EPI_M_DOTX) grid_reduce<3, BLOCK>(red, a.red, FinBiOmega{a.scal, a.status});
}

// The one-CTA tail of a reducing single-RHS step whose CTAs only stored their partials (ReduceBuf::split, reduce.cuh):
// same early-outs as the step itself, so that a stopped / predicated-off step leaves the scalars alone.
constexpr int FINISH_BLOCK = 1024;
template <typename R, int EPI>
__global__ void __launch_bounds__(FINISH_BLOCK) dslash_finish_kernel(const DslashArgs<R> a) {
  if (a.check_stop && (a.status[ST_STOP] != 0 || a.status[ST_BREAKDOWN] != 0)) return;
  if (a.run_if && a.status[a.run_if] == 0) return;
  if (EPI == EPI_M_NORM) reduce_finish<1, FINISH_BLOCK>(a.red, FinCgD{a.scal});
  if (EPI == EPI_M_CG) reduce_finish<1, FINISH_BLOCK>(a.red, FinCgCp{a.scal, a.status, a.iter, a.check_stop});
  if (EPI == EPI_M_CGREL) reduce_finish<1, FINISH_BLOCK>(a.red, FinRelCp{a.scal, a.status, a.iter, a.check_stop});
  if (EPI == EPI_M_DOTR0) reduce_finish<2, FINISH_BLOCK>(a.red, FinBiAlpha{a.scal, a.status});
  if (EPI == EPI_M_DOTX) reduce_finish<3, FINISH_BLOCK>(a.red, FinBiOmega{a.scal, a.status});
}

// ... and of a batched step: one CTA per right-hand side (blockIdx.x), each with its own partials, scalars and status.
template <typename R, int EPI>
__global__ void __launch_bounds__(FINISH_BLOCK) dslash_mrhs_finish_kernel(const DslashArgs<R> a) {
  const int rhs = blockIdx.x;
  double* const scal = a.scal + rhs * S_COUNT;
  int* const status = a.status + rhs * ST_COUNT;
  if (a.check_stop && (status[ST_STOP] != 0 || status[ST_BREAKDOWN] != 0)) return;
  if (EPI == EPI_M_NORM) reduce_finish<1, FINISH_BLOCK>(a.red.template for_rhs<1>(rhs), FinCgD{scal});
  if (EPI == EPI_M_CG) reduce_finish<1, FINISH_BLOCK>(a.red.template for_rhs<1>(rhs), FinCgCp{scal, status, a.iter, a.check_stop});
  if (EPI == EPI_M_DOTR0) reduce_finish<2, FINISH_BLOCK>(a.red.template for_rhs<2>(rhs), FinBiAlpha{scal, status});
  if (EPI == EPI_M_DOTX) reduce_finish<3, FINISH_BLOCK>(a.red.template for_rhs<3>(rhs), FinBiOmega{scal, status});
}

// ---- multi-RHS variant ------------------------------------------------------------------------------------------
// CTA = 32 consecutive target sites x NRB right-hand sides (threadIdx.y = right-hand side, one warp each).  All warps
// of a CTA need the same 8 links and the same clover block per site: the CTA stages them in shared memory once
// (cp.async), so gauge + clover cross L2/HBM once per CTA instead of once per right-hand side, occupy no registers
// while in flight, and the per-thread global loads are the spinors only -- per right-hand side the operator moves
// (120 + (144+16G)/nrhs) reals per site instead of 264+16G.
// Every right-hand side has its own scalar / status block and its own reduction (warp_grid_reduce), so the solves
// advance in lockstep but converge independently.
// Tuned on B200 at 48^3x48 with 12 sources (scripts/tune_mrhs.sh, profiles/r01_tune_mrhs.txt): the kernel is bound by
// latency x occupancy, so the best shapes are the ones that put the most warps on an SM without spilling --
// fp64: 6 sources/CTA x 2 CTAs/SM (12 warps, 170 registers) 1.71x per source over the single-RHS kernel
//       (4x2 at 255 registers: 1.34x; 4x4 at 128: 1.47x; 12x1: 1.69x);
// fp32: 12 sources/CTA x 2 CTAs/SM (24 warps, 85 registers) 1.53x (6x4: 1.45x; 4x4: 1.30x).
#ifndef B200_MRHS_NRB
#define B200_MRHS_NRB 6       // right-hand sides per CTA, fp64
#endif
#ifndef B200_MRHS_MINB
#define B200_MRHS_MINB 2      // min CTAs/SM for the register allocator, fp64
#endif
#ifndef B200_MRHS_NRB_F
#define B200_MRHS_NRB_F 12    // fp32
#endif
#ifndef B200_MRHS_MINB_F
#define B200_MRHS_MINB_F 2
#endif
template <typename R, int EPI, bool RECON12> struct MrhsSmem {
  static constexpr int NG = RECON12 ? 6 : 9;
  static constexpr int NL = 8 * NG;                                 // link slots
  static constexpr int NS = NL + (EPI == EPI_DSLASH ? 0 : 36);      // + clover slots
  // clover-late variant: only the links are staged up front; the 36 clover planes later overwrite link slots [0, 36),
  // which are dead once every warp is past hop HB = 36 / NG (the x and y links; with 12-number links the z links too)
  static constexpr bool CL = MrhsPrefetch<R>::clover_late && EPI != EPI_DSLASH;
  static constexpr int HB = 36 / NG;
  static constexpr int NSTAGE = CL ? NL : NS;                      // slots staged before the first hop
  static constexpr int CLOV0 = CL ? 0 : NL;                        // first clover slot
  static constexpr size_t bytes = (size_t)NSTAGE * 32 * sizeof(Cx<R>);
  // + `depth` 12-plane spinor buffers per right-hand side (warp) of the CTA
  static constexpr size_t total(int nrb) { return bytes + (MrhsPrefetch<R>::on ? (size_t)nrb * MrhsPrefetch<R>::depth * 12 * 32 * sizeof(Cx<R>) : 0); }
};
template <typename R, int EPI, bool RECON12, int NRB, int MODE = MODE_ASYM>
__global__ void __launch_bounds__(32 * NRB, (sizeof(R) == 4 ? B200_MRHS_MINB_F : B200_MRHS_MINB)) dslash_mrhs_kernel(const DslashArgs<R> a0, const LinkScale ls, int ngroups, const MrhsDiv dv) {
  typedef Cx<R> C;
  typedef MrhsSmem<R, EPI, RECON12> SM;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C* const sm = reinterpret_cast<C*>(smem_raw);
  // blockIdx.x = site_block * ngroups + grp (ngroups is 1, 2 or 3: cheaper as a compare chain than as a division)
  int site_block, grp;
  if (ngroups == 1) { site_block = blockIdx.x; grp = 0; }
  else if (ngroups == 2) { site_block = blockIdx.x >> 1; grp = blockIdx.x & 1; }
  else { site_block = blockIdx.x / ngroups; grp = blockIdx.x - site_block * ngroups; }
  const int rhs = grp * NRB + threadIdx.y;
  const int stride = a0.g.Vh;
  const int local = site_block * 32 + threadIdx.x;
  const bool active = local < a0.nsites;
  SiteCoord sc = {0, 0, 0, 0, 0};
  if (active) sc = mrhs_site<R>(a0, dv, local);
  const int idx = sc.idx;

  // ---- stage this CTA's operator data: the 8 links (and the clover block) of its 32 sites go to shared memory ONCE,
  // with asynchronous copies issued by all NRB warps; every right-hand side then reads them from there.  The slots
  // form groups of NG planes with one base pointer each (8 link directions, then the clover planes NG at a time); a
  // warp takes whole groups, so the choice of base pointer is a warp-uniform branch and the copies are a plain
  // unrolled "base + k * stride".
  const size_t gplane = (size_t)SM::NG * stride, gmu = 2 * gplane;
  if (active) {
    constexpr int NGRP = SM::NSTAGE / SM::NG;
    static_assert(NGRP * SM::NG == SM::NSTAGE, "slot groups");
#pragma unroll
    for (int j = 0; j < (NGRP + NRB - 1) / NRB; ++j) {
      const int gi = threadIdx.y + j * NRB;
      if (gi < NGRP) {
        const C* src;
        if (gi < 8) {
          const int mu = gi >> 1;
          if (gi & 1) src = a0.gauge + (size_t)(1 - a0.parity) * gplane + mu * gmu + backward_neighbour(a0.g, sc, a0.parity, mu);
          else src = a0.gauge + (size_t)a0.parity * gplane + mu * gmu + idx;
        } else {
          src = a0.clov + (size_t)(gi - 8) * gplane + idx;
        }
        C* const dst = sm + gi * (SM::NG * 32) + threadIdx.x;
#pragma unroll
        for (int k = 0; k < SM::NG; ++k) cp_async(dst + k * 32, src + (size_t)k * stride);
      }
    }
  }
  cp_async_commit();

  const bool have_rhs = rhs < a0.nrhs;
  constexpr bool PF = MrhsPrefetch<R>::on;
  constexpr int PD = MrhsPrefetch<R>::depth;
  C* const sp = sm + SM::NSTAGE * 32 + threadIdx.y * (PD * 12 * 32) + threadIdx.x;     // this lane's column of the warp's spinor buffers
  // the first hops' neighbour spinors are requested BEFORE the wait on the staged links: the fetches fly together
  if (PF && active && have_rhs) dslash_site_pf_begin<R>(a0.in + rhs * a0.fstride, a0.g, sc, a0.parity, sp, a0.pol.keep);
  else if (PF) { for (int k = 0; k < PD; ++k) cp_async_commit(); }

  DslashArgs<R> a = a0;
  if (have_rhs) {
    a.scal = a0.scal + rhs * S_COUNT; a.status = a0.status + rhs * ST_COUNT;
    a.in = a0.in + rhs * a0.fstride;
    if (a0.out) a.out = a0.out + rhs * a0.fstride;
    if (a0.x) a.x = a0.x + rhs * a0.fstride;
    if (a0.r) a.r = a0.r + rhs * a0.fstride;
    if (a0.r0) a.r0 = a0.r0 + rhs * a0.fstride;
    if (a0.ghost_fwd) { a.ghost_fwd = a0.ghost_fwd + rhs * a0.gstride; a.ghost_bwd = a0.ghost_bwd + rhs * a0.gstride; }
    if (a0.ghost_zfwd) { a.ghost_zfwd = a0.ghost_zfwd + rhs * a0.gstride_z; a.ghost_zbwd = a0.ghost_zbwd + rhs * a0.gstride_z; }
  }
  // a converged right-hand side (or an empty slot of the last group) only helps with the staging
  const bool work = have_rhs && !(a.check_stop && (a.status[ST_STOP] != 0 || a.status[ST_BREAKDOWN] != 0));
  if (PF) cp_async_wait_pending<PD>(); else cp_async_wait_all();      // the links + clover have landed (the spinors may still fly)
  __syncthreads();
  double red[3] = {0.0, 0.0, 0.0};
  const L2Policy pol = a.pol;
  constexpr bool XS = (EPI >= EPI_M) && PF;

  if (SM::CL) {
    // clover-late: hops 0 .. HB-1, barrier (the links of those hops are dead in every warp), every warp copies its share
    // of the clover block over them -- the copies join the commit group of the next hop's fetch and land while the
    // remaining hops run -- barrier, epilogue.  Warps without work walk through the same two barriers.
    C acc[12];
    if (work && active) dslash_site_pf<R, RECON12, 0, SM::HB>(acc, a, ls, sc, pol, sm + threadIdx.x, sp, XS ? a.x + idx : nullptr);
    __syncthreads();
    if (active) {
      constexpr int NCG = 36 / SM::NG;
#pragma unroll
      for (int j = 0; j < (NCG + NRB - 1) / NRB; ++j) {
        const int gi = threadIdx.y + j * NRB;
        if (gi < NCG) {
          const C* const src = a0.clov + (size_t)gi * gplane + idx;
          C* const dst = sm + gi * (SM::NG * 32) + threadIdx.x;
#pragma unroll
          for (int k = 0; k < SM::NG; ++k) cp_async(dst + k * 32, src + (size_t)k * stride);
        }
      }
    }
    if (work && active) dslash_site_pf<R, RECON12, SM::HB, 8>(acc, a, ls, sc, pol, sm + threadIdx.x, sp, XS ? a.x + idx : nullptr);
    else cp_async_commit();
    cp_async_wait_all();
    __syncthreads();
    if (!work) return;
    if (active) site_epilogue<R, EPI, true, MODE, XS>(acc, a, idx, stride, pol, red, sm + SM::CLOV0 * 32 + threadIdx.x, sp + (8 % PD) * (12 * 32));
  } else {
    if (!work) {
      cp_async_wait_all();
      return;
    }
    if (active) {
      C acc[12];
      if (PF) {
        dslash_site_pf<R, RECON12>(acc, a, ls, sc, pol, sm + threadIdx.x, sp, XS ? a.x + idx : nullptr);
        site_epilogue<R, EPI, true, MODE, XS>(acc, a, idx, stride, pol, red, sm + SM::CLOV0 * 32 + threadIdx.x, sp + (8 % PD) * (12 * 32));
      } else {
        dslash_site<R, RECON12, true>(acc, a, ls, idx, pol, sm + threadIdx.x, &sc);
        site_epilogue<R, EPI, true, MODE>(acc, a, idx, stride, pol, red, sm + SM::CLOV0 * 32 + threadIdx.x);
      }
    }
  }

  const int sb = site_block;   // block_offset of a split step is applied inside warp_grid_reduce
  if (a0.red.split) {             // big grids: partials only, dslash_mrhs_finish_kernel behind the step sums them
    if (EPI == EPI_M_NORM || EPI == EPI_M_CG) warp_partial_store<1>(red, a0.red, rhs, sb);
    if (EPI == EPI_M_DOTR0) warp_partial_store<2>(red, a0.red, rhs, sb);
    if (EPI == EPI_M_DOTX) warp_partial_store<3>(red, a0.red, rhs, sb);
    return;
  }
  if (EPI == EPI_M_NORM) warp_grid_reduce<1>(red, a0.red.template for_rhs<1>(rhs), sb, FinCgD{a.scal});
  if (EPI == EPI_M_CG) warp_grid_reduce<1>(red, a0.red.template for_rhs<1>(rhs), sb, FinCgCp{a.scal, a.status, a.iter, a.check_stop});
  if (EPI == EPI_M_DOTR0) warp_grid_reduce<2>(red, a0.red.template for_rhs<2>(rhs), sb, FinBiAlpha{a.scal, a.status});
  if (EPI == EPI_M_DOTX) warp_grid_reduce<3>(red, a0.red.template for_rhs<3>(rhs), sb, FinBiOmega{a.scal, a.status});
}

}  // namespace b200
