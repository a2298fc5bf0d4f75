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
