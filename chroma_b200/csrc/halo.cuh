// halo.cuh -- T-split multi-GPU support: half-spinor face exchange and cross-GPU reductions over NVLink
// peer memory (CUDA IPC), no host round trip and no library call inside the solver loop.
//
// Replaces the QMP face exchange of the reference's multi-node Dslash (tables_parscalar.h:35-113, 296-319;
// cpp_dslash_parscalar_64bit.cc:19-93) and the QMP global sums behind QDP++'s norm2/innerProduct.
// As in the reference, what crosses the cut is the spin-PROJECTED half spinor (12 reals per face site), and for
// the backward hop the SENDER multiplies by U^dagger (decomp_hvv, cpp_dslash_parscalar_utils_64bit.cc:61-110),
// so the hopping term needs no gauge ghost.
//
// Protocol per Dslash number n (all ranks issue the same sequence of Dslashes):
//   pack kernel   projects both time faces of the source field and STORES them straight into the neighbours'
//                 ghost buffers (slot n&1) through peer-mapped pointers; its last block then writes n into the
//                 neighbours' arrival flags (threadfence_system before the flag).
//   interior      dslash kernel on time slices 1 .. Lt-2 runs while the faces are in flight.
//   wait kernel   one thread spins on the two local arrival flags until both equal n.
//   boundary      dslash kernel on slices 0 and Lt-1, reading the ghost half spinors.
// Two slots are enough: a rank can only start packing Dslash n+2 after it finished the boundary of n+1, which
// needed the neighbour's pack n+1, which the neighbour issued after ITS boundary kernel of n had read slot n&1.
// Predicated Dslashes (run_if, used by the reliable-update solver) are skipped by ALL ranks or by none (the flag is
// derived from a global sum) and always come in groups of four, so executed Dslashes still alternate slots.
#pragma once
#include <vector>

#include "engine.cuh"
#include "dslash.cuh"
#include "clover_setup.cuh"

namespace b200 {

struct HaloLayout {
  size_t flags_off, seq_off, mailbox_off, ghost_off, gauge_ghost_off, total;
  size_t ghost_face;   // elements (complex) of one ghost face buffer: 6*S3h
};

template <typename R>
struct PackArgs {
  typedef Cx<R> C;
  const C* in;          // source field (parity src_par)
  const C* gauge;
  C* to_bwd;            // -t neighbour's ghost_fwd buffer of this slot (peer pointer)
  C* to_fwd;            // +t neighbour's ghost_bwd buffer of this slot (peer pointer)
  unsigned long long* flag_bwd;   // -t neighbour's arrival flag [slot][0]
  unsigned long long* flag_fwd;   // +t neighbour's arrival flag [slot][1]
  unsigned long long seq;
  unsigned int* ticket;
  const int* status;    // may be null
  const int* pred;      // may be null: status word that must be non-zero for this launch to run (predicated Dslash)
  Geom g;
  int src_par, isign, recon12;
  double scale_b;       // RECON12 only: aniso[3] * (bc_t if this rank owns the global last slice)
  int nrhs;             // batched Dslash: gridDim.y right-hand sides, fields fstride apart, ghost faces gstride apart
  size_t fstride, gstride;
};

template <typename R, bool RECON12>
__global__ void __launch_bounds__(128) pack_faces_kernel(const PackArgs<R> a) {
  typedef Cx<R> C;
  // A converged right-hand side sends nothing.  With a single right-hand side the whole Dslash (wait kernel included)
  // is skipped; in a batch the flags are still published so that the other right-hand sides can proceed.
  const int rhs = blockIdx.y;
  bool skip = false;
  if (a.status) { const int* st = a.status + rhs * ST_COUNT; skip = st[ST_STOP] != 0 || st[ST_BREAKDOWN] != 0; }
  if (a.nrhs == 1 && skip) return;
  if (a.pred && *a.pred == 0) return;
  const Geom& g = a.g;
  const int tid = blockIdx.x * 128 + threadIdx.x;
  const int stride = g.Vh, st = g.S3h;
  const R s = (R)a.isign;
  const L2Policy pol = make_l2_policy();
  const C* __restrict__ in = a.in + rhs * a.fstride;
  C* __restrict__ to_bwd = a.to_bwd + rhs * a.gstride;
  C* __restrict__ to_fwd = a.to_fwd + rhs * a.gstride;
  if (skip) {
  } else if (tid < st) {
    // face t = 0 -> forward-hop half spinor (1 - s g3) psi for the -t neighbour's slice Lt-1
    C h0[3], h1[3];
    load_project<R, 3>(h0, h1, in + tid, stride, -s, pol.keep);
#pragma unroll
    for (int c = 0; c < 3; ++c) { to_bwd[(size_t)c * st + tid] = h0[c]; to_bwd[(size_t)(3 + c) * st + tid] = h1[c]; }
  } else if (tid < 2 * st) {
    // face t = Lt-1 -> backward-hop half spinor U_t^dag (1 + s g3) psi for the +t neighbour's slice 0
    const int s3 = tid - st, idx = (g.Lt - 1) * st + s3;
    constexpr int NG = RECON12 ? 6 : 9;
    C h0[3], h1[3], U[9], r0[3], r1[3];
    load_project<R, 3>(h0, h1, in + idx, stride, s, pol.keep);
    load_link<R, RECON12>(U, a.gauge + ((size_t)(3 * 2 + a.src_par) * NG) * stride + idx, stride, pol.keep);
    if (RECON12) {
      const R sc = (R)a.scale_b;
#pragma unroll
      for (int c = 0; c < 3; ++c) { h0[c].x *= sc; h0[c].y *= sc; h1[c].x *= sc; h1[c].y *= sc; }
    }
    su3_mul<R, true>(r0, r1, U, h0, h1);
#pragma unroll
    for (int c = 0; c < 3; ++c) { to_fwd[(size_t)c * st + s3] = r0[c]; to_fwd[(size_t)(3 + c) * st + s3] = r1[c]; }
  }
  // last block publishes the arrival flags
  __threadfence_system();
  __shared__ bool is_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int t = atomicAdd(a.ticket, 1u);
    is_last = (t == gridDim.x * gridDim.y - 1);
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    __threadfence_system();
    *(volatile unsigned long long*)a.flag_bwd = a.seq;
    *(volatile unsigned long long*)a.flag_fwd = a.seq;
    *a.ticket = 0u;
    __threadfence_system();
  }
}

__global__ void wait_flags_kernel(const unsigned long long* f0, const unsigned long long* f1, unsigned long long seq, int* status, int check_stop, int run_if);

// One-time push of the boundary link slices needed by the field-strength (clover leaves reach x +/- t):
// face 0 of the receiver = slice t=-1 (sender's last slice), face 1 = slice t=Lt (sender's first slice).
template <typename R>
struct GaugePushArgs {
  CloverSetupArgs<R> cs;        // for fetch_link (original links incl. phases, no anisotropy)
  Cx<R>* to_fwd_face0;          // +t neighbour's gauge ghost, face 0
  Cx<R>* to_bwd_face1;          // -t neighbour's gauge ghost, face 1
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
class Halo {
 public:
  typedef Cx<R> C;
  int nranks = 1, rank = 0, fwd = 0, bwd = 0, device = 0;
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
    l.flags_off = 0;                                  // [2 slots][2 dirs] u64
    l.seq_off = 256;                                  // u64 reduction counters [MAX_RHS]
    l.mailbox_off = 512;                              // [MAX_RHS][2][8][8] doubles
    l.ghost_off = 512 + (size_t)MAX_RHS * MAILBOX_DOUBLES * sizeof(double);
    l.ghost_face = (size_t)6 * g.S3h;                 // one right-hand side of one face; a face buffer holds MAX_RHS of them
    l.gauge_ghost_off = l.ghost_off + 4 * l.ghost_face * MAX_RHS * sizeof(C);
    l.gauge_ghost_off = (l.gauge_ghost_off + 255) / 256 * 256;
    l.total = l.gauge_ghost_off + (size_t)2 * 4 * 2 * 9 * g.S3h * sizeof(C);
    return l;
  }

  int init(const Config& cfg, const Geom& g_, cudaStream_t s) {
    g = g_; stream = s; device = cfg.device;
    nranks = cfg.pgrid[3]; rank = cfg.pcoord[3];
    comm = cfg.comm;
    if (comm.size != nranks || comm.rank != rank || !comm.allgather || !comm.barrier) {
      set_error("b200_comm (rank %d/%d) does not match the T process grid (coord %d of %d)", comm.rank, comm.size, rank, nranks);
      return B200_ERR_COMM;
    }
    fwd = (rank + 1) % nranks; bwd = (rank + nranks - 1) % nranks;
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
    return B200_OK;
  }

  void destroy() {
    if (!arena) return;
    cudaDeviceSynchronize();
    if (comm.barrier) comm.barrier(comm.user);     // nobody may still be writing into a peer arena
    for (int r = 0; r < nranks; ++r) if (r != rank && peer[r]) cudaIpcCloseMemHandle(peer[r]);
    if (comm.barrier) comm.barrier(comm.user);
    cudaFree(arena); cudaFree(ticket);
    arena = nullptr;
  }

  C* ghost_ptr(char* base, int slot, int dir) const { return (C*)(base + lay.ghost_off) + (size_t)(slot * 2 + dir) * lay.ghost_face * MAX_RHS; }
  unsigned long long* flag_ptr(char* base, int slot, int dir) const { return (unsigned long long*)(base + lay.flags_off) + slot * 2 + dir; }
  const C* ghost_fwd() const { return ghost_ptr(arena, (int)(seq & 1), 0); }
  const C* ghost_bwd() const { return ghost_ptr(arena, (int)(seq & 1), 1); }
  C* gauge_ghost_of(char* base) const { return (C*)(base + lay.gauge_ghost_off); }
  const C* gauge_ghost() const { return gauge_ghost_of(arena); }

  int* status_dev = nullptr;             // engine status block, for timeouts
  PeerReduce peer_reduce() const {
    PeerReduce p;
    memset(&p, 0, sizeof(p));
    p.nranks = nranks; p.rank = rank; p.status = status_dev;
    if (nranks > 1) {
      p.seq = (unsigned long long*)(arena + lay.seq_off);
      for (int r = 0; r < nranks; ++r) p.mailbox[r] = (double*)(peer[r] + lay.mailbox_off);
    }
    return p;
  }

  // Pack + send both faces of `in` for the Dslash that targets `parity`.
  int start(const C* in, const C* gauge, int recon, const LinkScale& ls, int isign, int parity, const int* status, int run_if, int nrhs,
            size_t fstride, long long& launches) {
    ++seq;
    const int slot = (int)(seq & 1);
    PackArgs<R> a;
    a.in = in; a.gauge = gauge;
    a.to_bwd = ghost_ptr(peer[bwd], slot, 0);
    a.to_fwd = ghost_ptr(peer[fwd], slot, 1);
    a.flag_bwd = flag_ptr(peer[bwd], slot, 0);
    a.flag_fwd = flag_ptr(peer[fwd], slot, 1);
    a.seq = seq; a.ticket = ticket; a.status = status; a.pred = run_if ? status_dev + run_if : nullptr; a.g = g;
    a.src_par = 1 - parity; a.isign = isign; a.recon12 = recon == 12;
    a.scale_b = ls.aniso[3] * (ls.t_is_last ? (double)ls.bc_t : 1.0);
    a.nrhs = nrhs; a.fstride = fstride; a.gstride = lay.ghost_face;
    const dim3 blocks((2 * g.S3h + 127) / 128, nrhs);
    if (recon == 12) pack_faces_kernel<R, true><<<blocks, 128, 0, stream>>>(a);
    else pack_faces_kernel<R, false><<<blocks, 128, 0, stream>>>(a);
    ++launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("pack_faces launch failed: %s", cudaGetErrorString(e)); return B200_ERR_CUDA; }
    return B200_OK;
  }

  int wait(const int* status, int run_if, int nrhs, long long& launches) {
    const int slot = (int)(seq & 1);
    // a batch always waits: its flags are always published, and one right-hand side's stop flag says nothing about the others
    wait_flags_kernel<<<1, 1, 0, stream>>>(flag_ptr(arena, slot, 0), flag_ptr(arena, slot, 1), seq, status_dev, (status && nrhs == 1) ? 1 : 0, run_if);
    ++launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("wait_flags launch failed: %s", cudaGetErrorString(e)); return B200_ERR_CUDA; }
    return B200_OK;
  }

  int exchange_gauge_ghost(const CloverSetupArgs<R>& cs, long long& launches) {
    GaugePushArgs<R> a;
    a.cs = cs;
    a.to_fwd_face0 = gauge_ghost_of(peer[fwd]);
    a.to_bwd_face1 = gauge_ghost_of(peer[bwd]) + (size_t)4 * 2 * 9 * g.S3h;
    push_gauge_kernel<R><<<(4 * g.S3h + 127) / 128, 128, 0, stream>>>(a);
    ++launches;
    B200_CUDA(cudaGetLastError());
    B200_CUDA(cudaStreamSynchronize(stream));
    if (comm.barrier(comm.user) != 0) { set_error("barrier failed"); return B200_ERR_COMM; }
    return B200_OK;
  }
};

}  // namespace b200
