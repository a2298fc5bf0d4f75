// pack.cuh -- host-layout (QDP++ AoS) <-> device site-major SoA reordering, done ON THE GPU.
//
// Host arrays are copied raw (cudaMemcpyAsync, chunked through a staging buffer) and transposed here,
// so the host never touches a site loop.  Replaces qdp_pack_gauge (qdp_packer_nopad.cc:7-21), the
// anisotropy fold of CPPWilsonDslashD::create (lwldslash_w_cppd.cc:115-121) and the host packers of the
// QPhiX adapter (syssolver_linop_clover_qphix_w.h:166-187, 339-357).
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int PACK_SITES = 64;   // sites per block
constexpr int PACK_BLOCK = 256;

// plane -> offset (in reals) of its (re,im) pair inside one site's AoS record
struct MapIdentity { __host__ __device__ int operator()(int plane) const { return 2 * plane; } };
// PrimitiveClovTriang: diag[2][6] | offd[2][15][2]  ->  planes block*18 + {3 diag pairs, 15 offd}
struct MapClover {
  __host__ __device__ int operator()(int plane) const {
    const int b = plane / 18, q = plane % 18;
    return q < 3 ? 6 * b + 2 * q : 12 + 30 * b + 2 * (q - 3);
  }
};

// src: H[nsites][NR] AoS (site-contiguous).  dst plane p element s -> dst[p*stride + dst_off + s].
// Only planes [0, NPL) are written (NPL < NR/2 drops trailing planes: 12-real gauge compression).
template <typename H, typename R, int NR, int NPL, typename Map>
__global__ void __launch_bounds__(PACK_BLOCK) aos_to_soa_kernel(const H* __restrict__ src, Cx<R>* __restrict__ dst, int nsites,
                                                               size_t stride, size_t dst_off, Map map, double scale) {
  constexpr int PAD = NR + 1;
  __shared__ H tile[PACK_SITES * PAD];
  const int base = blockIdx.x * PACK_SITES;
  const int ns = min(PACK_SITES, nsites - base);
  const H* s = src + (size_t)base * NR;
  for (int i = threadIdx.x; i < ns * NR; i += PACK_BLOCK) tile[(i / NR) * PAD + (i % NR)] = s[i];
  __syncthreads();
  for (int i = threadIdx.x; i < NPL * PACK_SITES; i += PACK_BLOCK) {
    const int pl = i / PACK_SITES, site = i % PACK_SITES;
    if (site < ns) {
      const int off = map(pl);
      dst[(size_t)pl * stride + dst_off + base + site] =
          mk<R>((R)(scale * (double)tile[site * PAD + off]), (R)(scale * (double)tile[site * PAD + off + 1]));
    }
  }
}

// 12-real link compression keeps rows 0 and 1 and rebuilds row 2 = conj(row0 x row1) in the kernels.  That is only right
// for SU(3) links; the one non-unit factor the engine carries separately is the -1 of an antiperiodic T boundary on the
// last global time slice.  Links with any other phase (spatially antiperiodic or twisted fermion boundaries, U(3) links)
// would be reconstructed wrongly, so the upload checks every link: |conj(row0 x row1) - sgn * row2| with sgn = -1 for
// sites in [flip_lo, flip_hi) (index within this parity field), +1 elsewhere; the maximum is kept as the bit pattern of a
// non-negative double (ordered like an unsigned integer).
template <typename H>
__global__ void __launch_bounds__(256) recon12_check_kernel(const H* __restrict__ aos, int nsites, size_t site_off, size_t flip_lo, size_t flip_hi,
                                                          unsigned long long* maxdev) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  double dev = 0.0;
  if (i < nsites) {
    const H* u = aos + (size_t)i * 18;
    const size_t site = site_off + i;
    const double sg = (site >= flip_lo && site < flip_hi) ? -1.0 : 1.0;
    for (int c = 0; c < 3; ++c) {
      const int c1 = (c + 1) % 3, c2 = (c + 2) % 3;
      const double ar = u[2 * c1], ai = u[2 * c1 + 1], br = u[6 + 2 * c2], bi = u[6 + 2 * c2 + 1];
      const double cr = u[2 * c2], ci = u[2 * c2 + 1], dr = u[6 + 2 * c1], di = u[6 + 2 * c1 + 1];
      const double tr = (ar * br - ai * bi) - (cr * dr - ci * di), ti = (ar * bi + ai * br) - (cr * di + ci * dr);
      const double er = tr - sg * (double)u[12 + 2 * c], ei = -ti - sg * (double)u[12 + 2 * c + 1];
      dev = fmax(dev, fmax(fabs(er), fabs(ei)));
    }
  }
  for (int o = 16; o > 0; o >>= 1) dev = fmax(dev, __shfl_xor_sync(0xffffffffu, dev, o));
  if ((threadIdx.x & 31) == 0 && dev > 0.0) atomicMax(maxdev, (unsigned long long)__double_as_longlong(dev));
}

template <typename H, typename R, int NR, int NPL, typename Map>
__global__ void __launch_bounds__(PACK_BLOCK) soa_to_aos_kernel(H* __restrict__ dst, const Cx<R>* __restrict__ src, int nsites,
                                                               size_t stride, size_t src_off, Map map) {
  constexpr int PAD = NR + 1;
  __shared__ H tile[PACK_SITES * PAD];
  const int base = blockIdx.x * PACK_SITES;
  const int ns = min(PACK_SITES, nsites - base);
  for (int i = threadIdx.x; i < NPL * PACK_SITES; i += PACK_BLOCK) {
    const int pl = i / PACK_SITES, site = i % PACK_SITES;
    if (site < ns) {
      const Cx<R> v = src[(size_t)pl * stride + src_off + base + site];
      const int off = map(pl);
      tile[site * PAD + off] = (H)v.x; tile[site * PAD + off + 1] = (H)v.y;
    }
  }
  __syncthreads();
  H* d = dst + (size_t)base * NR;
  for (int i = threadIdx.x; i < ns * NR; i += PACK_BLOCK) d[i] = tile[(i / NR) * PAD + (i % NR)];
}

}  // namespace b200
