// reduce.cuh -- deterministic block reduction + "last block finalises" pattern used by every fused
// solver kernel.  Each block writes one partial per reduced quantity; the block that draws the last
// ticket sums the partials in a fixed order (so the result does not depend on block scheduling),
// optionally combines across GPUs through NVLink peer mailboxes, and hands the totals to a finaliser
// functor that derives the solver scalars (alpha, beta, ...) ON THE DEVICE -- the host never sees them.
#pragma once
#include "common.cuh"

namespace b200 {

// Cross-GPU reduction mailboxes (peer-mapped).  See comm.cuh for how they are wired up.
constexpr int MAILBOX_DOUBLES = 2 * 8 * 8;   // per right-hand side: [slot 2][src rank 8][4 values + tag + pad]
struct PeerReduce {
  int nranks;           // 1 => single GPU, nothing to do
  int rank;
  unsigned long long* seq;  // device counters [MAX_RHS]: cross-GPU reductions done so far for each right-hand side
  double* mailbox[8];       // mailbox[r] = rank r's mailbox base (peer pointer), layout [rhs][slot 2][src rank 8][4 values + seq]
  int* status;              // device status block (ST_BREAKDOWN = 91 on timeout)
};

struct ReduceBuf {
  double* partial;       // [N][total_blocks]
  unsigned int* ticket;  // zero between kernels
  int block_offset;      // first partial slot of this launch (a step may be split into several launches)
  int total_blocks;      // blocks over all launches of the step
  PeerReduce peer;
  // The view of right-hand side `rhs` of a batched step reducing N quantities: own partials, ticket, mailbox, status.
  template <int N>
  __device__ __forceinline__ ReduceBuf for_rhs(int rhs) const {
    ReduceBuf r = *this;
    r.partial += (size_t)rhs * N * total_blocks;
    r.ticket += rhs;
    if (r.peer.nranks > 1) {
      r.peer.seq += rhs;
      for (int i = 0; i < r.peer.nranks; ++i) r.peer.mailbox[i] += (size_t)rhs * MAILBOX_DOUBLES;
    }
    if (r.peer.status) r.peer.status += rhs * ST_COUNT;
    return r;
  }
};

template <int BLOCK>
__device__ __forceinline__ double block_sum(double v, double* smem /* [BLOCK/32] */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) smem[w] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < BLOCK / 32; ++i) s += smem[i];   // every thread sums in the same fixed order
  return s;
}

// Combine `vals[N]` (N<=4) across ranks in fixed rank order.  Executed by ONE thread of the last block.
// Each rank writes its values + a sequence tag into every peer's mailbox (slot = seq&1), then polls its
// own mailbox until all N ranks' tags match.  Two slots suffice because a rank cannot start reduction
// s+2 before every peer has finished reading reduction s (they all needed this rank's s+1 contribution).
template <int N>
__device__ __forceinline__ void peer_allreduce(const PeerReduce& pr, double vals[N]) {
  if (pr.nranks <= 1) return;
  const unsigned long long s = *pr.seq + 1ull;
  const int slot = (int)(s & 1ull);
  for (int dst = 0; dst < pr.nranks; ++dst) {
    volatile double* mb = pr.mailbox[dst] + ((size_t)slot * 8 + pr.rank) * 8;
    for (int k = 0; k < N; ++k) mb[k] = vals[k];
  }
  __threadfence_system();
  for (int dst = 0; dst < pr.nranks; ++dst) {
    volatile unsigned long long* tag =
        (volatile unsigned long long*)(pr.mailbox[dst] + ((size_t)slot * 8 + pr.rank) * 8 + 4);
    *tag = s;
  }
  __threadfence_system();
  double tot[N];
  for (int k = 0; k < N; ++k) tot[k] = 0.0;
  for (int src = 0; src < pr.nranks; ++src) {
    volatile double* mb = pr.mailbox[pr.rank] + ((size_t)slot * 8 + src) * 8;
    volatile unsigned long long* tag = (volatile unsigned long long*)(mb + 4);
    const long long t0 = clock64();
    while (*tag != s) {
      if (clock64() - t0 > PEER_SPIN_CYCLES) { if (pr.status) pr.status[ST_BREAKDOWN] = 91; break; }
    }
    __threadfence_system();
    for (int k = 0; k < N; ++k) tot[k] += mb[k];
  }
  for (int k = 0; k < N; ++k) vals[k] = tot[k];
  *pr.seq = s;
}

// v[N]: this thread's contributions.  fin(tot) runs in exactly one thread of the whole step.
template <int N, int BLOCK, typename Fin>
__device__ __forceinline__ void grid_reduce(double v[N], const ReduceBuf& rb, Fin fin) {
  __shared__ double smem[BLOCK / 32];
  __shared__ bool is_last;
  double s[N];
#pragma unroll
  for (int k = 0; k < N; ++k) s[k] = block_sum<BLOCK>(v[k], smem);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < N; ++k) rb.partial[(size_t)k * rb.total_blocks + rb.block_offset + blockIdx.x] = s[k];
    __threadfence();
    unsigned int t = atomicAdd(rb.ticket, 1u);
    is_last = (t == (unsigned int)rb.total_blocks - 1u);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    double tot[N];
#pragma unroll
    for (int k = 0; k < N; ++k) {
      double acc = 0.0;
      for (int b = threadIdx.x; b < rb.total_blocks; b += BLOCK) acc += __ldcg(rb.partial + (size_t)k * rb.total_blocks + b);
      tot[k] = block_sum<BLOCK>(acc, smem);
    }
    if (threadIdx.x == 0) {
      peer_allreduce<N>(rb.peer, tot);
      fin(tot);
      *rb.ticket = 0u;
      __threadfence();
    }
  }
}

// Split variant for the big single-RHS Dslash grids (B200_SPLIT_REDUCE): in grid_reduce every CTA must wait for its
// ticket atomic to come back before it may exit (it has to learn whether it is the last one), and with one CTA per SM
// that round trip -- tens of thousands of CTAs hammering one address -- is dead time on the SM.  Here the CTAs only
// store their partials and leave; a one-CTA kernel launched behind the step sums them in the SAME fixed order, combines
// across GPUs and runs the finaliser.
template <int N, int BLOCK>
__device__ __forceinline__ void block_partials(double v[N], const ReduceBuf& rb) {
  __shared__ double smem[BLOCK / 32];
  double s[N];
#pragma unroll
  for (int k = 0; k < N; ++k) s[k] = block_sum<BLOCK>(v[k], smem);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < N; ++k) rb.partial[(size_t)k * rb.total_blocks + rb.block_offset + blockIdx.x] = s[k];
  }
}
template <int N, int BLOCK, typename Fin>
__device__ __forceinline__ void finish_partials(const ReduceBuf& rb, Fin fin) {
  __shared__ double smem[BLOCK / 32];
  double tot[N];
#pragma unroll
  for (int k = 0; k < N; ++k) {
    double acc = 0.0;
    for (int b = threadIdx.x; b < rb.total_blocks; b += BLOCK) acc += __ldcg(rb.partial + (size_t)k * rb.total_blocks + b);
    tot[k] = block_sum<BLOCK>(acc, smem);
  }
  if (threadIdx.x == 0) {
    peer_allreduce<N>(rb.peer, tot);
    fin(tot);
    __threadfence();
  }
}

// Warp-synchronous variant for the multi-RHS Dslash kernels, where one WARP (32 consecutive sites of one right-hand
// side) is the reduction unit: shuffle tree -> one partial per (rhs, site block) -> the warp that draws the last
// ticket of its right-hand side sums that right-hand side's partials (lane-strided, then the same shuffle tree: a
// fixed order) and lane 0 runs the finaliser.  rb must already be the for_rhs() view.  No __syncthreads: warps of
// one CTA belong to different right-hand sides and may have returned early (converged).
template <int N, typename Fin>
__device__ __forceinline__ void warp_grid_reduce(double v[N], const ReduceBuf& rb, int site_block, Fin fin) {
  const int lane = threadIdx.x & 31;
  double s[N];
#pragma unroll
  for (int k = 0; k < N; ++k) {
    double x = v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    s[k] = x;
  }
  unsigned int t = 0;
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < N; ++k) rb.partial[(size_t)k * rb.total_blocks + rb.block_offset + site_block] = s[k];
    __threadfence();
    t = atomicAdd(rb.ticket, 1u);
  }
  t = __shfl_sync(0xffffffffu, t, 0);
  if (t != (unsigned int)rb.total_blocks - 1u) return;
  __threadfence();
  double tot[N];
#pragma unroll
  for (int k = 0; k < N; ++k) {
    double acc = 0.0;
    for (int b = lane; b < rb.total_blocks; b += 32) acc += __ldcg(rb.partial + (size_t)k * rb.total_blocks + b);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    tot[k] = acc;
  }
  if (lane == 0) {
    peer_allreduce<N>(rb.peer, tot);
    fin(tot);
    *rb.ticket = 0u;
    __threadfence();
  }
}

}  // namespace b200
