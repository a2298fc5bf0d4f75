// halo.cuh -- multi-GPU support (T split, then T x Z): half-spinor face exchange and cross-GPU reductions over NVLink
// peer memory (CUDA IPC), no host round trip and no library call inside the solver loop.
//
// Replaces the QMP face exchange of the reference's multi-node Dslash (tables_parscalar.h:35-113, 296-319;
// cpp_dslash_parscalar_64bit.cc:19-93) and the QMP global sums behind QDP++'s norm2/innerProduct.
// As in the reference, what crosses the cut is the spin-PROJECTED half spinor (12 reals per face site), and for
// the backward hop the SENDER multiplies by U^dagger (decomp_hvv, cpp_dslash_parscalar_utils_64bit.cc:61-110),
// so the hopping term needs no gauge ghost.
//
// Protocol per Dslash number n (all ranks issue the same sequence of Dslashes).  A single-RHS Dslash is ONE launch
// (dslash_halo_kernel below) whose CTA ranges play the roles; a batched Dslash still issues them as separate kernels:
//   pack          projects the faces of the source field (t = 0, Lt-1; on a T x Z grid also z = 0, Lz-1) and STORES
//                 them straight into the neighbours' ghost buffers (slot n&1) through peer-mapped pointers; the last
//                 pack CTA then writes n into the neighbours' arrival flags (CTA barrier, one system fence, flag).
//   interior      the sites with no ghost dependence run while the faces are in flight.
//   wait          thread 0 of every boundary CTA (batched: a one-thread kernel) polls the local arrival flags (2 or 4)
//                 until all equal n.
//   boundary      the boundary slices / planes, reading the ghost half spinors through the coherent path.
// Two slots are enough: a rank can only start packing Dslash n+2 after it finished the boundary of n+1, which
// needed the neighbour's pack n+1, which the neighbour issued after ITS boundary CTAs of n had read slot n&1.
// Predicated Dslashes (run_if, used by the reliable-update solver) are skipped by ALL ranks or by none (the flag is
// derived from a global sum) and always come in groups of four, so executed Dslashes still alternate slots.
#pragma once
#include <vector>

#include "engine.cuh"
#include "dslash.cuh"
#include "clover_setup.cuh"

namespace b200 {

struct HaloLayout {
  size_t flags_off, seq_off, mailbox_off, ghost_off, gauge_ghost_off, gauge_ghost_z_off, total;
  size_t face_t, face_z;   // elements (complex) of one right-hand side of one ghost face: 6*S3h (T), 6*SZh (Z)
  size_t slot_elems;       // elements of one slot = both T faces + both Z faces, MAX_RHS right-hand sides each
};

// Faces of the pack kernel / ghost buffers of the receiver, in this order everywhere:
//   0: my t=0 plane     -> (1 - s g3) psi               -> the -t neighbour's "T forward" ghost  (its sites at t = Lt-1)
//   1: my t=Lt-1 plane  -> U_t^dag (1 + s g3) psi        -> the +t neighbour's "T backward" ghost (its sites at t = 0)
//   2: my z=0 plane     -> (1 - s g2) psi               -> the -z neighbour's "Z forward" ghost
//   3: my z=Lz-1 plane  -> U_z^dag (1 + s g2) psi        -> the +z neighbour's "Z backward" ghost
template <typename R>
struct PackArgs {
  typedef Cx<R> C;
  const C* in;          // source field (parity src_par)
  const C* gauge;
  C* to[4];             // destination ghost buffer of each face in the neighbour's arena (peer pointer), null if not split
  unsigned long long* flag[4];   // that neighbour's arrival flag for the face
  unsigned long long seq;
  unsigned int* ticket;
  const int* status;    // may be null
  const int* pred;      // may be null: status word that must be non-zero for this launch to run (predicated Dslash)
  Geom g;
  int src_par, isign, recon12;
  double scale_b[2];    // RECON12 only: factor of the backward faces: aniso[3] * (bc_t if this rank owns the global last slice); aniso[2]
  int nrhs;             // batched Dslash: gridDim.y right-hand sides, fields fstride apart, ghost faces gstride apart
  size_t fstride, gstride[2];   // gstride: T faces, Z faces
};

// forward-hop face: project only
template <typename R, int MU>
__device__ __forceinline__ void pack_project(Cx<R>* __restrict__ dst, int f, int fs, const Cx<R>* __restrict__ src, int stride, R sg, const L2Policy& pol) {
  Cx<R> h0[3], h1[3];
  load_project<R, MU>(h0, h1, src, stride, sg, pol.keep);
#pragma unroll
  for (int c = 0; c < 3; ++c) { dst[(size_t)c * fs + f] = h0[c]; dst[(size_t)(3 + c) * fs + f] = h1[c]; }
}
// backward-hop face: project and multiply by U^dagger (the sender owns the link)
template <typename R, int MU, bool RECON12>
__device__ __forceinline__ void pack_project_mul(Cx<R>* __restrict__ dst, int f, int fs, const Cx<R>* __restrict__ src, const Cx<R>* __restrict__ link,
                                                 int stride, R sg, R scale, const L2Policy& pol) {
  Cx<R> h0[3], h1[3], U[9], r0[3], r1[3];
  load_project<R, MU>(h0, h1, src, stride, sg, pol.keep);
  load_link<R, RECON12>(U, link, stride, pol.keep);
  if (RECON12) {
#pragma unroll
    for (int c = 0; c < 3; ++c) { h0[c].x *= scale; h0[c].y *= scale; h1[c].x *= scale; h1[c].y *= scale; }
  }
  su3_mul<R, true>(r0, r1, U, h0, h1);
#pragma unroll
  for (int c = 0; c < 3; ++c) { dst[(size_t)c * fs + f] = r0[c]; dst[(size_t)(3 + c) * fs + f] = r1[c]; }
}

// Body of the face pack for CTA `bx` of `nbx` (x RHS `rhs` of `nby`), BLOCK threads each.  The CTA strides over the
// face sites (nbx may be smaller than the number of BLOCK-sized pieces: the fused kernel uses one pack CTA per SM so that
// interior CTAs start in the other slot of every SM at once and the pack overlaps them instead of preceding them).
template <typename R, bool RECON12, int BLOCK>
__device__ __forceinline__ void pack_faces_body(const PackArgs<R>& a, int bx, int nbx, int rhs, int nby) {
  typedef Cx<R> C;
  // A converged right-hand side sends nothing.  With a single right-hand side the whole Dslash (wait included)
  // is skipped; in a batch the flags are still published so that the other right-hand sides can proceed.
  bool skip = false;
  if (a.status) { const int* st = a.status + rhs * ST_COUNT; skip = st[ST_STOP] != 0 || st[ST_BREAKDOWN] != 0; }
  if (a.nrhs == 1 && skip) return;
  if (a.pred && *a.pred == 0) return;
  const Geom& g = a.g;
  const int stride = g.Vh, st = g.S3h, sz = g.SZh, row = g.Lxh * g.Ly;
  const R s = (R)a.isign;
  const L2Policy pol = make_l2_policy();
  constexpr int NG = RECON12 ? 6 : 9;
  const C* __restrict__ in = a.in + rhs * a.fstride;
  const int nT = g.tsplit ? 2 * st : 0, nZ = g.zsplit ? 2 * sz : 0;
  if (!skip) {
    for (int tid = bx * BLOCK + threadIdx.x; tid < nT + nZ; tid += nbx * BLOCK) {
      if (tid < nT) {
        if (tid < st) {
          pack_project<R, 3>(a.to[0] + rhs * a.gstride[0], tid, st, in + tid, stride, -s, pol);
        } else {
          const int f = tid - st, idx = (g.Lt - 1) * st + f;
          pack_project_mul<R, 3, RECON12>(a.to[1] + rhs * a.gstride[0], f, st, in + idx, a.gauge + ((size_t)(3 * 2 + a.src_par) * NG) * stride + idx,
                                          stride, s, (R)a.scale_b[0], pol);
        }
      } else {
        const int tz = tid - nT;
        const bool back = tz >= sz;
        const int f = back ? tz - sz : tz;                 // face index (t*Ly + y)*Lxh + xh
        const int t = f / row, w = f - t * row;
        const int idx = t * st + (back ? (g.Lz - 1) * row : 0) + w;
        if (!back) pack_project<R, 2>(a.to[2] + rhs * a.gstride[1], f, sz, in + idx, stride, -s, pol);
        else pack_project_mul<R, 2, RECON12>(a.to[3] + rhs * a.gstride[1], f, sz, in + idx, a.gauge + ((size_t)(2 * 2 + a.src_par) * NG) * stride + idx,
                                             stride, s, (R)a.scale_b[1], pol);
      }
    }
  }
  // The last CTA publishes the arrival flags.  One system-scope fence per CTA, by the thread that draws the ticket, after
  // a CTA barrier: the barrier orders every thread's peer stores before thread 0's fence, and fences are cumulative, so
  // the stores are visible to the peer GPU before the flag written behind the last ticket.
  __shared__ bool is_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    unsigned int t = atomicAdd(a.ticket, 1u);
    is_last = (t == (unsigned int)(nbx * nby) - 1u);
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    __threadfence_system();
#pragma unroll
    for (int f = 0; f < 4; ++f) if (a.flag[f]) *(volatile unsigned long long*)a.flag[f] = a.seq;
    *a.ticket = 0u;
    __threadfence_system();
  }
}
template <typename R, bool RECON12>
__global__ void __launch_bounds__(128) pack_faces_kernel(const PackArgs<R> a) {
  pack_faces_body<R, RECON12, 128>(a, blockIdx.x, gridDim.x, blockIdx.y, gridDim.y);
}

// ---- the whole split-lattice Dslash in ONE launch (single right-hand side) ------------------------------------------
// Round 1 issued pack -> interior -> wait<<<1,1>>> -> boundary as four kernels on one stream: three extra kernel tails
// per Dslash, the pack overlapped with nothing, and the boundary could not start before the last interior CTA had
// left (17 launches per CG iteration on a split lattice).  Here the roles are CTA ranges of one grid, in dispatch order:
//   [0, n_pack)                 project the faces, store them into the neighbours' ghost buffers, last one publishes seq
//   [n_pack, n_pack + n_int)    interior sites (no ghost dependence): run while the faces are in flight
//   the rest                    boundary sites: thread 0 spins on the local arrival flags, then the CTA proceeds
// The boundary CTAs are the last of the grid, so they only occupy SM slots once every pack / interior CTA has been
// dispatched; what they wait for is the NEIGHBOURS' pack CTAs, the first CTAs of the neighbours' launch of the same
// Dslash, which need nothing from this rank's launch.  A CG iteration on a split lattice is 5 launches, as on one GPU.
template <typename R>
struct HaloFuse {
  PackArgs<R> pack;
  const unsigned long long* wait[4];   // local arrival flags of the faces in use (null = not split)
  unsigned long long seq;
  long long spin;                      // spin budget (SM cycles) before a lost peer raises status 90
  int n_pack, n_int;                   // CTAs of the first two roles
  int n_int_sites;                     // sites of box[0] (the interior); boundary CTAs enumerate box[1..]
};
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
template <typename R, int EPI, bool RECON12, int BLOCK, int MODE = MODE_ASYM>
__global__ void __launch_bounds__(BLOCK, (sizeof(R) == 4 ? B200_DSLASH_MINBLOCKS_F : B200_DSLASH_MINBLOCKS))
dslash_halo_kernel(const DslashArgs<R> a, const LinkScale ls, const HaloFuse<R> h, const MrhsDiv dv) {
  typedef Cx<R> C;
  if (a.check_stop && (a.status[ST_STOP] != 0 || a.status[ST_BREAKDOWN] != 0)) return;
  if (a.run_if && a.status[a.run_if] == 0) return;
  int b = blockIdx.x;
  if (b < h.n_pack) { pack_faces_body<R, RECON12, BLOCK>(h.pack, b, h.n_pack, 0, 1); return; }
  b -= h.n_pack;
  const bool boundary = b >= h.n_int;
  if (boundary) {
    if (threadIdx.x == 0) {
      const long long t0 = clock64();
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        if (!h.wait[f]) continue;
        while (ld_acquire_sys(h.wait[f]) < h.seq) {
          if (clock64() - t0 > h.spin) { a.status[ST_BREAKDOWN] = 90; break; }
          __nanosleep(64);
        }
      }
    }
    __syncthreads();
  }
  const int stride = a.g.Vh;
  const int local = boundary ? (b - h.n_int) * BLOCK + threadIdx.x : b * BLOCK + threadIdx.x;
  const bool active = boundary ? local < a.nsites - h.n_int_sites : local < h.n_int_sites;
  double red[3] = {0.0, 0.0, 0.0};
  if (active) {
    const SiteCoord sc = boundary ? site_coord(a.g, dv, launch_site_from<R>(a, local, 1)) : mrhs_site<R>(a, dv, local);
    const int idx = sc.idx;
    const L2Policy pol = a.pol;
    C acc[12];
    dslash_site<R, RECON12, false>(acc, a, ls, idx, pol, nullptr, &sc);
    site_epilogue<R, EPI, false, MODE>(acc, a, idx, stride, pol, red);
  }
  if (EPI == EPI_M_NORM) grid_reduce<1, BLOCK>(red, a.red, FinCgD{a.scal}, b);
  if (EPI == EPI_M_CG) grid_reduce<1, BLOCK>(red, a.red, FinCgCp{a.scal, a.status, a.iter, a.check_stop}, b);
  if (EPI == EPI_M_CGREL) grid_reduce<1, BLOCK>(red, a.red, FinRelCp{a.scal, a.status, a.iter, a.check_stop}, b);
  if (EPI == EPI_M_DOTR0) grid_reduce<2, BLOCK>(red, a.red, FinBiAlpha{a.scal, a.status}, b);
  if (EPI == EPI_M_DOTX) grid_reduce<3, BLOCK>(red, a.red, FinBiOmega{a.scal, a.status}, b);
}

struct WaitFlags { const unsigned long long* f[4]; };   // local arrival flags of the faces in use (null = not split)
__global__ void wait_flags_kernel(WaitFlags w, unsigned long long seq, int* status, int check_stop, int run_if, long long spin);

// One-time push of the boundary links the field strength needs (clover leaves reach x +/- mu +/- nu).
// T faces: face 0 of the receiver = slice t=-1 (sender's last slice), face 1 = slice t=Lt (sender's first slice).
// Z faces (T x Z grids) are EXTENDED in t to t = -1 .. Lt, so that they carry the corner links x -/+ t -/+ z: they are
// pushed AFTER the T faces have arrived, and the rows t = -1 / Lt are read from the sender's own T ghost.
template <typename R>
struct GaugePushArgs {
  CloverSetupArgs<R> cs;        // for fetch_link (original links incl. phases, no anisotropy)
  Cx<R>* to_fwd_face0;          // +mu neighbour's gauge ghost, face 0
  Cx<R>* to_bwd_face1;          // -mu neighbour's gauge ghost, face 1
};
template <typename R>
__global__ void __launch_bounds__(128) push_gauge_kernel(const GaugePushArgs<R> a) {
  const Geom& g = a.cs.g;
  const int tid = blockIdx.x * 128 + threadIdx.x;
  const int st = g.S3h;
  if (tid >= 2 * 2 * st) return;
  const int face = tid / (2 * st), rem = tid % (2 * st), par = rem / st, s3 = rem % st;
  const int t = face == 0 ? g.Lt - 1 : 0;            // my slice that goes out
  int q = s3;
  const int xh = q % g.Lxh; q /= g.Lxh;
  const int y = q % g.Ly; const int z = q / g.Ly;
  const int x = 2 * xh + ((y + z + t + par) & 1);
  Cx<R>* dst = face == 0 ? a.to_fwd_face0 : a.to_bwd_face1;
  for (int mu = 0; mu < 4; ++mu) {
    Z U[9];
    fetch_link<R>(U, a.cs, mu, x, y, z, t);
    Cx<R>* p = dst + ((size_t)(mu * 2 + par) * 9) * st + s3;
    for (int k = 0; k < 9; ++k) p[(size_t)k * st] = mk<R>((R)U[k].x, (R)U[k].y);
  }
}
template <typename R>
__global__ void __launch_bounds__(128) push_gauge_z_kernel(const GaugePushArgs<R> a) {
  const Geom& g = a.cs.g;
  const int tid = blockIdx.x * 128 + threadIdx.x;
  const int sze = g.Lxh * g.Ly * (g.Lt + 2);         // one extended Z face of one parity
  if (tid >= 2 * 2 * sze) return;
  const int face = tid / (2 * sze), rem = tid % (2 * sze), par = rem / sze, f = rem % sze;
  const int z = face == 0 ? g.Lz - 1 : 0;            // my plane that goes out
  int q = f;
  const int xh = q % g.Lxh; q /= g.Lxh;
  const int y = q % g.Ly; const int t = q / g.Ly - 1;   // -1 .. Lt
  const int x = 2 * xh + ((y + z + t + par + 2) & 1);
  Cx<R>* dst = face == 0 ? a.to_fwd_face0 : a.to_bwd_face1;
  for (int mu = 0; mu < 4; ++mu) {
    Z U[9];
    fetch_link<R>(U, a.cs, mu, x, y, z, t);
    Cx<R>* p = dst + ((size_t)(mu * 2 + par) * 9) * sze + f;
    for (int k = 0; k < 9; ++k) p[(size_t)k * sze] = mk<R>((R)U[k].x, (R)U[k].y);
  }
}

template <typename R>
class Halo {
 public:
  typedef Cx<R> C;
  int nranks = 1, rank = 0, device = 0;
  int nbr[4] = {0, 0, 0, 0};             // rank that receives face f of the pack kernel: -t, +t, -z, +z neighbour
  Geom g{};
  cudaStream_t stream = nullptr;
  char* arena = nullptr;                 // local
  std::vector<char*> peer;               // peer[r] = base of rank r's arena as mapped here (peer[rank] = arena)
  HaloLayout lay{};
  unsigned long long seq = 0;            // Dslash counter (host side, identical on all ranks)
  unsigned int* ticket = nullptr;
  b200_comm comm{};

  static HaloLayout layout(const Geom& g) {
    HaloLayout l;
    l.flags_off = 0;                                  // [2 slots][4 faces] u64
    l.seq_off = 256;                                  // u64 reduction counters [MAX_RHS]
    l.mailbox_off = 512;                              // [MAX_RHS][2][8][8] doubles
    l.ghost_off = 512 + (size_t)MAX_RHS * MAILBOX_DOUBLES * sizeof(double);
    l.face_t = g.tsplit ? (size_t)6 * g.S3h : 0;      // one right-hand side of one face; a face buffer holds MAX_RHS of them
    l.face_z = g.zsplit ? (size_t)6 * g.SZh : 0;
    l.slot_elems = 2 * (l.face_t + l.face_z) * MAX_RHS;
    l.gauge_ghost_off = l.ghost_off + 2 * l.slot_elems * sizeof(C);
    l.gauge_ghost_off = (l.gauge_ghost_off + 255) / 256 * 256;
    l.gauge_ghost_z_off = l.gauge_ghost_off + (g.tsplit ? (size_t)2 * 4 * 2 * 9 * g.S3h * sizeof(C) : 0);
    l.gauge_ghost_z_off = (l.gauge_ghost_z_off + 255) / 256 * 256;
    l.total = l.gauge_ghost_z_off + (g.zsplit ? (size_t)2 * 4 * 2 * 9 * g.Lxh * g.Ly * (g.Lt + 2) * sizeof(C) : 0) + 256;
    return l;
  }

  // Rank of grid coordinate (pz, pt): x fastest, like QMP's logical topology (here px = py = 0).
  static int rank_of(const Config& cfg, int pz, int pt) {
    const int Pz = cfg.pgrid[2], Pt = cfg.pgrid[3];
    return ((pt + Pt) % Pt) * Pz + (pz + Pz) % Pz;
  }

  int init(const Config& cfg, const Geom& g_, cudaStream_t s) {
    g = g_; stream = s; device = cfg.device;
    nranks = cfg.pgrid[2] * cfg.pgrid[3]; rank = rank_of(cfg, cfg.pcoord[2], cfg.pcoord[3]);
    comm = cfg.comm;
    if (comm.size != nranks || comm.rank != rank || !comm.allgather || !comm.barrier) {
      set_error("b200_comm (rank %d/%d) does not match the process grid: expected rank pt*Pz+pz = %d of %d", comm.rank, comm.size, rank, nranks);
      return B200_ERR_COMM;
    }
    nbr[0] = rank_of(cfg, cfg.pcoord[2], cfg.pcoord[3] - 1); nbr[1] = rank_of(cfg, cfg.pcoord[2], cfg.pcoord[3] + 1);
    nbr[2] = rank_of(cfg, cfg.pcoord[2] - 1, cfg.pcoord[3]); nbr[3] = rank_of(cfg, cfg.pcoord[2] + 1, cfg.pcoord[3]);
    if (const char* e = getenv("B200_PEER_TIMEOUT_S")) spin_cycles = (long long)(atof(e) * 2.0e9);
    lay = layout(g);
    B200_CUDA(cudaMalloc(&arena, lay.total));
    B200_CUDA(cudaMemset(arena, 0, lay.total));
    B200_CUDA(cudaMalloc(&ticket, sizeof(unsigned int)));
    B200_CUDA(cudaMemset(ticket, 0, sizeof(unsigned int)));
    B200_CUDA(cudaDeviceSynchronize());
    cudaIpcMemHandle_t mine;
    B200_CUDA(cudaIpcGetMemHandle(&mine, arena));
    std::vector<cudaIpcMemHandle_t> all(nranks);
    if (comm.allgather(comm.user, &mine, all.data(), sizeof(mine)) != 0) { set_error("allgather of IPC handles failed"); return B200_ERR_COMM; }
    peer.assign(nranks, nullptr);
    for (int r = 0; r < nranks; ++r) {
      if (r == rank) { peer[r] = arena; continue; }
      void* p = nullptr;
      cudaError_t e = cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) { set_error("cudaIpcOpenMemHandle(rank %d) failed: %s (are all ranks on one NVLink node?)", r, cudaGetErrorString(e)); return B200_ERR_COMM; }
      peer[r] = (char*)p;
    }
    if (comm.barrier(comm.user) != 0) { set_error("barrier failed"); return B200_ERR_COMM; }
    ready = true;
    return B200_OK;
  }

  bool ready = false;   // init() completed on this rank (and therefore, after its last barrier, on every rank)
  void destroy() {
    if (!arena) return;
    cudaDeviceSynchronize();
    // the barriers pair with those of the other ranks' destroy(); a rank whose init() failed half way must not wait for
    // peers that never got here (they would be waiting in init's own collective)
    if (ready && comm.barrier) comm.barrier(comm.user);     // nobody may still be writing into a peer arena
    for (int r = 0; r < (int)peer.size(); ++r) if (r != rank && peer[r]) cudaIpcCloseMemHandle(peer[r]);
    if (ready && comm.barrier) comm.barrier(comm.user);
    ready = false;
    cudaFree(arena); cudaFree(ticket);
    arena = nullptr;
  }

  // Ghost buffer that RECEIVES face f (see PackArgs) in slot `slot` of the arena at `base`.  Within a slot: the two T
  // faces, then the two Z faces.  The receiver reads face 0 as its "T forward" ghost, face 1 as "T backward", etc.
  C* ghost_ptr(char* base, int slot, int f) const {
    C* p = (C*)(base + lay.ghost_off) + (size_t)slot * lay.slot_elems;
    return f < 2 ? p + (size_t)f * lay.face_t * MAX_RHS : p + 2 * lay.face_t * MAX_RHS + (size_t)(f - 2) * lay.face_z * MAX_RHS;
  }
  unsigned long long* flag_ptr(char* base, int slot, int f) const { return (unsigned long long*)(base + lay.flags_off) + slot * 4 + f; }
  bool face_on(int f) const { return f < 2 ? g.tsplit != 0 : g.zsplit != 0; }
  const C* ghost(int f) const { return face_on(f) ? ghost_ptr(arena, (int)(seq & 1), f) : nullptr; }
  C* gauge_ghost_of(char* base) const { return (C*)(base + lay.gauge_ghost_off); }
  C* gauge_ghost_z_of(char* base) const { return (C*)(base + lay.gauge_ghost_z_off); }
  const C* gauge_ghost() const { return g.tsplit ? gauge_ghost_of(arena) : nullptr; }
  const C* gauge_ghost_z() const { return g.zsplit ? gauge_ghost_z_of(arena) : nullptr; }

  int* status_dev = nullptr;             // engine status block, for timeouts
  PeerReduce peer_reduce() const {
    PeerReduce p;
    memset(&p, 0, sizeof(p));
    p.nranks = nranks; p.rank = rank; p.status = status_dev; p.spin = spin_cycles;
    if (nranks > 1) {
      p.seq = (unsigned long long*)(arena + lay.seq_off);
      for (int r = 0; r < nranks; ++r) p.mailbox[r] = (double*)(peer[r] + lay.mailbox_off);
    }
    return p;
  }

  long long spin_cycles = PEER_SPIN_CYCLES;   // budget of every peer wait (SM clock cycles); B200_PEER_TIMEOUT_S overrides
  int pack_threads() const { return (g.tsplit ? 2 * g.S3h : 0) + (g.zsplit ? 2 * g.SZh : 0); }
  const unsigned long long* local_flag(int f) const { return face_on(f) ? flag_ptr(arena, (int)(seq & 1), f) : nullptr; }

  // Start Dslash number ++seq: fill the arguments of its face pack (source `in`, target `parity`).
  void prepare(PackArgs<R>& a, const C* in, const C* gauge, int recon, const LinkScale& ls, int isign, int parity, int nrhs, size_t fstride) {
    ++seq;
    const int slot = (int)(seq & 1);
    a.in = in; a.gauge = gauge;
    for (int f = 0; f < 4; ++f) {
      a.to[f] = face_on(f) ? ghost_ptr(peer[nbr[f]], slot, f) : nullptr;
      a.flag[f] = face_on(f) ? flag_ptr(peer[nbr[f]], slot, f) : nullptr;
    }
    a.seq = seq; a.ticket = ticket; a.status = nullptr; a.pred = nullptr; a.g = g;
    a.src_par = 1 - parity; a.isign = isign; a.recon12 = recon == 12;
    a.scale_b[0] = ls.aniso[3] * (ls.t_is_last ? (double)ls.bc_t : 1.0);
    a.scale_b[1] = ls.aniso[2];
    a.nrhs = nrhs; a.fstride = fstride; a.gstride[0] = lay.face_t; a.gstride[1] = lay.face_z;
  }

  // Pack + send all faces of `in` for the Dslash that targets `parity` (stand-alone pack kernel: batched right-hand sides).
  int start(const C* in, const C* gauge, int recon, const LinkScale& ls, int isign, int parity, const int* status, int run_if, int nrhs,
            size_t fstride, long long& launches) {
    PackArgs<R> a;
    prepare(a, in, gauge, recon, ls, isign, parity, nrhs, fstride);
    a.status = status; a.pred = run_if ? status_dev + run_if : nullptr;
    const dim3 blocks((pack_threads() + 127) / 128, nrhs);
    if (recon == 12) pack_faces_kernel<R, true><<<blocks, 128, 0, stream>>>(a);
    else pack_faces_kernel<R, false><<<blocks, 128, 0, stream>>>(a);
    ++launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("pack_faces launch failed: %s", cudaGetErrorString(e)); return B200_ERR_CUDA; }
    return B200_OK;
  }

  int wait(const int* status, int run_if, int nrhs, long long& launches) {
    const int slot = (int)(seq & 1);
    WaitFlags w;
    for (int f = 0; f < 4; ++f) w.f[f] = face_on(f) ? flag_ptr(arena, slot, f) : nullptr;
    // a batch always waits: its flags are always published, and one right-hand side's stop flag says nothing about the others
    wait_flags_kernel<<<1, 1, 0, stream>>>(w, seq, status_dev, (status && nrhs == 1) ? 1 : 0, run_if, spin_cycles);
    ++launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("wait_flags launch failed: %s", cudaGetErrorString(e)); return B200_ERR_CUDA; }
    return B200_OK;
  }

  // cs.ghost_links / cs.ghost_links_z must already point at this rank's gauge ghosts.
  int exchange_gauge_ghost(const CloverSetupArgs<R>& cs, long long& launches) {
    GaugePushArgs<R> a;
    a.cs = cs;
    if (g.tsplit) {
      a.to_fwd_face0 = gauge_ghost_of(peer[nbr[1]]);
      a.to_bwd_face1 = gauge_ghost_of(peer[nbr[0]]) + (size_t)4 * 2 * 9 * g.S3h;
      push_gauge_kernel<R><<<(4 * g.S3h + 127) / 128, 128, 0, stream>>>(a);
      ++launches;
      B200_CUDA(cudaGetLastError());
      B200_CUDA(cudaStreamSynchronize(stream));
      if (comm.barrier(comm.user) != 0) { set_error("barrier failed"); return B200_ERR_COMM; }
    }
    if (g.zsplit) {
      const size_t sze = (size_t)g.Lxh * g.Ly * (g.Lt + 2);
      a.to_fwd_face0 = gauge_ghost_z_of(peer[nbr[3]]);
      a.to_bwd_face1 = gauge_ghost_z_of(peer[nbr[2]]) + (size_t)4 * 2 * 9 * sze;
      push_gauge_z_kernel<R><<<(int)((4 * sze + 127) / 128), 128, 0, stream>>>(a);
      ++launches;
      B200_CUDA(cudaGetLastError());
      B200_CUDA(cudaStreamSynchronize(stream));
      if (comm.barrier(comm.user) != 0) { set_error("barrier failed"); return B200_ERR_COMM; }
    }
    return B200_OK;
  }
};

}  // namespace b200
