// common.cuh -- shared types of the B200 clover engine (device layout, geometry, small complex algebra).
//
// Device layout ("site-major SoA"): every field is a set of PLANES of complex numbers, one plane
// per internal index, each plane holding one checkerboard's sites in QDP++ cb2 order
//   idx = ((t*Lz+z)*Ly+y)*(Lx/2) + x/2
// so that consecutive threads (= consecutive idx) load consecutive 16-byte (fp64) complex numbers:
//   fermion  C[12][Vh]              plane = spin*3+colour
//   gauge    C[4][2][9][Vh]         [mu][parity][row*3+col]   (aniso factor folded in, BC phases as given)
//   clover   C[2][36][Vh]           [parity][block*18 + {3 diag pairs, 15 offd}]
// Vh = local checkerboard volume.  Plane stride == Vh.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

template <typename R> struct CT;
template <> struct CT<double> { typedef double2 type; };
template <> struct CT<float> { typedef float2 type; };
template <typename R> using Cx = typename CT<R>::type;

template <typename R> __host__ __device__ __forceinline__ Cx<R> mk(R x, R y) { Cx<R> r; r.x = x; r.y = y; return r; }

__device__ __forceinline__ double2 ldg(const double2* p) { return __ldg(p); }
__device__ __forceinline__ float2 ldg(const float2* p) { return __ldg(p); }

// ---- L2 residency control -------------------------------------------------------------------------
// One time slice of gauge + clover + output streams ~117 MB through the 126 MB L2 at 48^3 (fp64) while the
// neighbour spinors of that slice (10.6 MB) are re-read up to 8 times over three slices.  Streams are therefore
// loaded/stored with an evict-first policy (and no L1 allocation), neighbour spinors with evict-last, so that the
// re-reads are served by L2 instead of HBM (ncu: 13% excess DRAM reads without this).
// sm_100 only takes the direct .L2::evict_* qualifiers on 256-bit accesses; 128/64-bit ones need a createpolicy
// descriptor + .L2::cache_hint.
struct L2Policy { uint64_t keep, stream; };
__device__ __forceinline__ L2Policy make_l2_policy() {
  L2Policy p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p.keep));
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p.stream));
  return p;
}
__device__ __forceinline__ double2 ld_keep(const double2* p, uint64_t pol) {
  double2 v;
  asm("ld.global.nc.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ float2 ld_keep(const float2* p, uint64_t pol) {
  float2 v;
  asm("ld.global.nc.L2::cache_hint.v2.f32 {%0,%1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ double2 ld_stream(const double2* p, uint64_t pol) {
  double2 v;
  asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ float2 ld_stream(const float2* p, uint64_t pol) {
  float2 v;
  asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f32 {%0,%1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(pol));
  return v;
}
// Multi-RHS kernels: the 12 right-hand sides of a CTA read the SAME links / clover blocks, so those loads allocate in L1
// (plain ld.global.nc, default L2 policy) and the first warp's miss serves the other eleven; the spinors, which are
// private to one right-hand side, bypass L1 so they cannot evict the shared operator data.
__device__ __forceinline__ double2 ld_op(const double2* p) { return __ldg(p); }
__device__ __forceinline__ float2 ld_op(const float2* p) { return __ldg(p); }
__device__ __forceinline__ double2 ld_keep_nol1(const double2* p, uint64_t pol) {
  double2 v;
  asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ float2 ld_keep_nol1(const float2* p, uint64_t pol) {
  float2 v;
  asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f32 {%0,%1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(pol));
  return v;
}
// Asynchronous global -> shared copies of one complex number (the multi-RHS kernels stage links and clover blocks in
// shared memory).  16-byte copies bypass L1 (.cg); 8-byte ones only exist as .ca.
__device__ __forceinline__ void cp_async(double2* smem_dst, const double2* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async(float2* smem_dst, const float2* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
// ... with an L2 residency hint (the batched kernels prefetch the next hop's neighbour spinor into shared memory).
// L1 = true allocates the line in L1 as well (.ca); 16-byte copies may bypass it (.cg), 8-byte ones cannot.
template <bool L1>
__device__ __forceinline__ void cp_async_hint(double2* smem_dst, const double2* gsrc, uint64_t pol) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  if (L1) asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "l"(pol) : "memory");
  else asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "l"(pol) : "memory");
}
template <bool L1>
__device__ __forceinline__ void cp_async_hint(float2* smem_dst, const float2* gsrc, uint64_t pol) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 8, %2;" ::"r"(d), "l"(gsrc), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// read-modify-write streams (the CG residual) must not use the non-coherent path
__device__ __forceinline__ double2 ld_stream_rw(const double2* p, uint64_t pol) {
  double2 v;
  asm("ld.global.L1::no_allocate.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ float2 ld_stream_rw(const float2* p, uint64_t pol) {
  float2 v;
  asm("ld.global.L1::no_allocate.L2::cache_hint.v2.f32 {%0,%1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void st_stream(double2* p, double2 v, uint64_t pol) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.v2.f64 [%0], {%1,%2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_stream(float2* p, float2 v, uint64_t pol) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.v2.f32 [%0], {%1,%2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
}

// complex helpers (all on Cx<R>)
template <typename C> __device__ __forceinline__ C cadd(C a, C b) { a.x += b.x; a.y += b.y; return a; }
template <typename C> __device__ __forceinline__ C csub(C a, C b) { a.x -= b.x; a.y -= b.y; return a; }
template <typename C> __device__ __forceinline__ C cmul(C a, C b) { C r; r.x = a.x*b.x - a.y*b.y; r.y = a.x*b.y + a.y*b.x; return r; }
// acc += a*b
template <typename C> __device__ __forceinline__ void cmac(C& acc, C a, C b) {
  acc.x += a.x*b.x; acc.x -= a.y*b.y; acc.y += a.x*b.y; acc.y += a.y*b.x;
}
// acc += conj(a)*b
template <typename C> __device__ __forceinline__ void cmac_conj(C& acc, C a, C b) {
  acc.x += a.x*b.x; acc.x += a.y*b.y; acc.y += a.x*b.y; acc.y -= a.y*b.x;
}

struct Geom {
  int Lxh, Ly, Lz, Lt;  // local extents (x halved by checkerboarding)
  int Vh;               // local checkerboard volume = plane stride
  int S3h;              // sites per time slice per checkerboard = Lxh*Ly*Lz
  int tsplit;           // 1 if T is split across ranks (ghost faces instead of wrap-around in t)
  int zsplit;           // 1 if Z is split across ranks as well (T x Z process grid)
  int SZh;              // sites per z-plane per checkerboard = Lxh*Ly*Lt (one Z face)
};

// A box of target sites of one launch: t in [t0, t0+nt), z in [z0, z0+nz), every x and y.  A Dslash launch covers up
// to four boxes (the whole lattice; the interior of a split lattice; its boundary slices / planes).
struct SiteBox { int t0, nt, z0, nz; };
__host__ __device__ __forceinline__ int box_count(const Geom& g, const SiteBox& b) { return g.Lxh * g.Ly * b.nz * b.nt; }
// cb2 index of the `local`-th site of a box (local < box_count)
__host__ __device__ __forceinline__ int box_site(const Geom& g, const SiteBox& b, int local) {
  if (b.nz == g.Lz) return b.t0 * g.S3h + local;                 // whole time slices are contiguous in cb2 order
  const int row = g.Lxh * g.Ly;
  const int w = local % row, q = local / row;
  const int zz = q % b.nz, tt = q / b.nz;
  return ((b.t0 + tt) * g.Lz + b.z0 + zz) * row + w;
}

// Division by a launch-invariant divisor without the ~20-instruction integer-division sequence: q = (n * mul) >> shift
// with shift = 28 + ceil(log2 d), mul = ceil(2^shift / d) is exact for 0 <= n < 2^28, 1 <= d <= 2^28 (the error term
// n * (mul*d - 2^shift) stays below 2^shift).  Site counts of one rank are far below 2^28 (64^3 x 128 / 2 = 2^24).
struct FastDiv { unsigned mul, shift; };
inline FastDiv make_fastdiv(int d) {
  if (d < 1) d = 1;
  unsigned s = 0;
  while ((1ull << s) < (unsigned long long)d) ++s;
  FastDiv f;
  f.shift = 28 + s;
  f.mul = (unsigned)(((1ull << f.shift) + (unsigned long long)d - 1) / (unsigned long long)d);
  return f;
}
__host__ __device__ __forceinline__ int fast_div(int n, const FastDiv f) { return (int)(((unsigned long long)(unsigned)n * f.mul) >> f.shift); }

// Right-hand sides solved in lockstep by the batched (multi-RHS) kernels: the 12 spin-colour sources of a propagator
// (quarkprop4_w.cc:70-117).  Every right-hand side owns one ScalarSlot block and one StatusSlot block.
constexpr int MAX_RHS = 12;

// Device-resident scalars of the solvers (double, one block of S_COUNT per right-hand side).
enum ScalarSlot {
  S_RSDSQ = 0,   // stopping threshold |r|^2 <= / < rsd_sq
  S_C, S_D, S_CP, S_A, S_B,                       // CG: c=|r_{k-1}|^2, d=|Mp|^2, cp=|r_k|^2, a=c/d, b=cp/c
  S_RHO_RE, S_RHO_IM, S_RHOP_RE, S_RHOP_IM,       // BiCGStab rho, rho_prev
  S_ALPHA_RE, S_ALPHA_IM, S_OMEGA_RE, S_OMEGA_IM, S_BETA_RE, S_BETA_IM,
  S_RNORM,                                        // BiCGStab |r|^2
  S_TMP0, S_TMP1, S_TMP2, S_TMP3,                 // generic reduction results (norm2 / inner)
  S_R0NORM, S_MAXRX, S_MAXRR, S_DELTA,            // reliable updates (reliable_cg.cc:68-73,115-121)
  S_COUNT = 32
};
// Integer status block
enum StatusSlot {
  ST_STOP = 0,      // 0 while iterating; else the iteration at which the recurrence residual converged
  ST_BREAKDOWN,     // BiCGStab breakdown code (1 rho=0, 2 <r0|v>=0, 3 |t|=0); 90/91 = peer wait timed out (halo / reduction)
  ST_UPD_R,         // reliable updates: this iteration replaces the residual with the fp64 one (updateR, reliable_cg.cc:120)
  ST_UPD_X,         // ... and folds the accumulated fp32 solution into psi (updateX, :119)
  ST_NUPD,          // number of residual replacements done
  ST_COUNT = 8
};

// Default spin-wait budget for peer flags (SM clock cycles, ~60 s; B200_PEER_TIMEOUT_S overrides): a lost peer must not
// hang the GPU forever, but ranks may legitimately be seconds apart (lazy module loading, I/O on one rank).
constexpr long long PEER_SPIN_CYCLES = 120000000000LL;

}  // namespace b200
