// reduce.cuh -- deterministic grid reductions used by every fused solver kernel.  Each block writes one partial per
// reduced quantity; the partials are summed in a fixed two-level order (so the result does not depend on block
// scheduling) -- by the block that draws the last ticket (small grids) or by a one-CTA finish kernel (the big Dslash
// grids, ReduceBuf::split) --, optionally combined across GPUs through NVLink peer mailboxes (peer_allreduce), and the
// totals go to a finaliser functor that derives the solver scalars (alpha, beta, ...) ON THE DEVICE -- the host never
// sees them.
#pragma once
#include "common.cuh"

namespace b200 {

// Cross-GPU reduction mailboxes (peer-mapped; halo.cuh wires them up).
constexpr int MAILBOX_DOUBLES = 2 * 8 * 8;   // per right-hand side: [slot 2][src rank 8][8 words: 4 values x (lo, hi) data+tag words]
struct PeerReduce {
  int nranks;           // 1 => single GPU, nothing to do
  int rank;
  unsigned long long* seq;  // device counters [MAX_RHS]: cross-GPU reductions done so far for each right-hand side
  double* mailbox[8];       // mailbox[r] = rank r's mailbox base (peer pointer), layout [rhs][slot 2][src rank 8][4 values + seq]
  int* status;              // device status block (ST_BREAKDOWN = 91 on timeout)
  long long spin;           // spin budget in SM cycles
};

// Reductions over more than RED_FLAT_MAX blocks are summed in TWO levels: blocks form groups of RED_GROUP consecutive
// partial slots with one ticket per group; the block that draws the last ticket of its group sums the group (one
// warp, two loads per lane) and then draws the step's ticket; the last group finisher sums the group sums.  With a flat
// ticket the last of the 41 472 CTAs of a 48^3x96 Dslash summed all partials alone -- 324 dependent L2 round trips per
// thread, ~8 % of the kernel (round-1 ncu: EPI_M_NORM 2.05 ms against 1.89 ms for the same traffic).  Order of
// summation is fixed in both levels, so results stay bitwise reproducible.
constexpr int RED_GROUP = 64;
constexpr int RED_FLAT_MAX = 2048;
struct ReduceBuf {
  double* partial;       // [N][total_blocks]
  unsigned int* ticket;  // zero between kernels
  double* gpartial;      // [N][ngroups] group sums (two-level reductions)
  unsigned int* gticket; // [ngroups], zero between kernels
  int split;             // != 0: the blocks only store their partials; reduce_finish (a one-CTA kernel behind the step) sums them
  int block_offset;      // first partial slot of this launch (a step may be split into several launches)
  int total_blocks;      // blocks over all launches of the step
  PeerReduce peer;
  __host__ __device__ __forceinline__ int ngroups() const { return (total_blocks + RED_GROUP - 1) / RED_GROUP; }
  // The view of right-hand side `rhs` of a batched step reducing N quantities: own partials, tickets, mailbox, status.
  template <int N>
  __device__ __forceinline__ ReduceBuf for_rhs(int rhs) const {
    ReduceBuf r = *this;
    r.partial += (size_t)rhs * N * total_blocks;
    r.ticket += rhs;
    r.gpartial += (size_t)rhs * N * ngroups();
    r.gticket += (size_t)rhs * ngroups();
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

// Combine `vals[N]` (N<=4) across ranks in fixed rank order.  Executed by ONE WARP (all 32 lanes call it with the same
// vals; lane r talks to rank r); on return every lane holds the totals.
// Protocol (NCCL's "LL" idea): every 8-byte word that crosses NVLink carries 4 bytes of payload and a 4-byte tag, and
// 8-byte stores are never torn, so a word is either old or complete -- no fence between data and flag, no fence at all.
// A double travels as two such words.  Rank `me` writes its N values into slot (seq&1, me) of EVERY rank's mailbox (lane =
// destination: eight posted stores in flight at once), then lane r polls slot (seq&1, r) of the LOCAL mailbox until both
// tags of every value match, and the totals are formed in rank order with shuffles, so all ranks get bit-identical sums.
// Round 1 did this in one thread with two system fences and serial polling: ~20 us per reduction at 8 GPUs.
// Two slots suffice because a rank cannot start reduction s+2 before every peer has finished reading reduction s (they
// all needed this rank's s+1 contribution).
template <int N>
__device__ __forceinline__ void peer_allreduce(const PeerReduce& pr, double vals[N]) {
  if (pr.nranks <= 1) return;
  const int lane = threadIdx.x & 31;
  const unsigned long long s = *pr.seq + 1ull;
  const int slot = (int)(s & 1ull);
  const unsigned long long tag = ((s & 0x7fffffffull) | 0x80000000ull) << 32;
  if (lane < pr.nranks) {
    volatile unsigned long long* mb = (volatile unsigned long long*)(pr.mailbox[lane] + ((size_t)slot * 8 + pr.rank) * 8);
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const unsigned long long bits = (unsigned long long)__double_as_longlong(vals[k]);
      mb[2 * k] = tag | (bits & 0xffffffffull);
      mb[2 * k + 1] = tag | (bits >> 32);
    }
  }
  double mine[N];
#pragma unroll
  for (int k = 0; k < N; ++k) mine[k] = 0.0;
  if (lane < pr.nranks) {
    const volatile unsigned long long* mb = (const volatile unsigned long long*)(pr.mailbox[pr.rank] + ((size_t)slot * 8 + lane) * 8);
    const long long t0 = clock64();
#pragma unroll
    for (int k = 0; k < N; ++k) {
      unsigned long long lo, hi;
      while ((((lo = mb[2 * k]) ^ tag) >> 32) != 0ull || (((hi = mb[2 * k + 1]) ^ tag) >> 32) != 0ull) {
        if (clock64() - t0 > pr.spin) { if (pr.status) pr.status[ST_BREAKDOWN] = 91; lo = hi = 0ull; break; }
      }
      mine[k] = __longlong_as_double((long long)((lo & 0xffffffffull) | (hi << 32)));
    }
  }
#pragma unroll
  for (int k = 0; k < N; ++k) {
    double tot = 0.0;
    for (int src = 0; src < pr.nranks; ++src) tot += __shfl_sync(0xffffffffu, mine[k], src);
    vals[k] = tot;
  }
  if (lane == 0) *pr.seq = s;
  __syncwarp();
}

// Draw a ticket.  The release semantics of the atomic order this thread's partial-sum stores before the ticket becomes
// visible; a plain __threadfence() here would be a full fence and ptxas implements that as MEMBAR + CCTL.IVALL, i.e. every
// CTA would invalidate its SM's whole L1 on the way out (B300_MICROARCH.md, "L1D flush trigger") and stall on the
// membar while the SM's other CTA is starved of issue slots.  Only the finishing CTA needs the acquire side.
#ifndef B200_RED_RELEASE
#define B200_RED_RELEASE 1
#endif
__device__ __forceinline__ unsigned int draw_ticket(unsigned int* p) {
#if B200_RED_RELEASE
  unsigned int old;
  asm volatile("atom.add.release.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(p) : "memory");
  return old;
#else
  __threadfence();
  return atomicAdd(p, 1u);
#endif
}

// Sum of n doubles at p by the calling thread group: thread `tid` of `nthr` takes p[tid], p[tid+nthr], ... four
// independent loads at a time (a fixed order, and four L2 round trips in flight instead of one).
__device__ __forceinline__ double strided_sum(const double* p, int n, int tid, int nthr) {
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  int b = tid;
  for (; b + 3 * nthr < n; b += 4 * nthr) {
    const double x0 = __ldcg(p + b), x1 = __ldcg(p + b + nthr), x2 = __ldcg(p + b + 2 * nthr), x3 = __ldcg(p + b + 3 * nthr);
    a0 += x0; a1 += x1; a2 += x2; a3 += x3;
  }
  for (; b < n; b += nthr) a0 += __ldcg(p + b);
  return (a0 + a1) + (a2 + a3);
}
// Sum of the <= RED_GROUP partials of group `grp` by one warp (fixed order: lane, lane+32, then the shuffle tree).
__device__ __forceinline__ double group_sum(const double* p, int gsize, int lane) {
  double x = (lane < gsize ? __ldcg(p + lane) : 0.0) + (lane + 32 < gsize ? __ldcg(p + lane + 32) : 0.0);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

// v[N]: this thread's contributions.  fin(tot) runs in exactly one thread of the whole step.  blk = this block's
// index among the reducing blocks of its launch (blockIdx.x unless the launch has other CTAs in front).
#ifndef B200_RED_PROBE
#define B200_RED_PROBE 0   // measurement probes (wrong sums!): 1 = accumulate only, 2 = + block sums and partial stores, no ticket
#endif
template <int N, int BLOCK, typename Fin>
__device__ __forceinline__ void grid_reduce(double v[N], const ReduceBuf& rb, Fin fin, int blk = -1) {
#if B200_RED_PROBE == 1
  if (v[0] == 1.2345e-300) rb.partial[0] = v[N - 1];
  return;
#endif
  __shared__ double smem[BLOCK / 32];
  __shared__ int role1, role2;    // role1: this block drew the last ticket of its group (flat: of the step); role2: ... of the step
  if (blk < 0) blk = blockIdx.x;
  blk += rb.block_offset;
  const bool flat = rb.total_blocks <= RED_FLAT_MAX;
  const int grp = blk / RED_GROUP, ngrp = rb.ngroups();
  const int gsize = min(RED_GROUP, rb.total_blocks - grp * RED_GROUP);
  double s[N];
#pragma unroll
  for (int k = 0; k < N; ++k) s[k] = block_sum<BLOCK>(v[k], smem);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < N; ++k) rb.partial[(size_t)k * rb.total_blocks + blk] = s[k];
  }
  if (rb.split) return;            // uniform over the grid: no ticket, nobody waits (see reduce_finish)
  if (threadIdx.x == 0) {
#if B200_RED_PROBE == 2
    role1 = 0;
#else
    if (flat) role1 = draw_ticket(rb.ticket) == (unsigned int)rb.total_blocks - 1u;
    else role1 = draw_ticket(rb.gticket + grp) == (unsigned int)gsize - 1u;
#endif
  }
  __syncthreads();
  if (!role1) return;
  if (!flat) {
    if (threadIdx.x < 32) {
      __threadfence();
      double gs[N];
#pragma unroll
      for (int k = 0; k < N; ++k) gs[k] = group_sum(rb.partial + (size_t)k * rb.total_blocks + grp * RED_GROUP, gsize, threadIdx.x);
      if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < N; ++k) rb.gpartial[(size_t)k * ngrp + grp] = gs[k];
        rb.gticket[grp] = 0u;
        role2 = draw_ticket(rb.ticket) == (unsigned int)ngrp - 1u;
      }
    }
    __syncthreads();
    if (!role2) return;
  }
  __threadfence();
  double tot[N];
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const double acc = flat ? strided_sum(rb.partial + (size_t)k * rb.total_blocks, rb.total_blocks, threadIdx.x, BLOCK)
                            : strided_sum(rb.gpartial + (size_t)k * ngrp, ngrp, threadIdx.x, BLOCK);
    tot[k] = block_sum<BLOCK>(acc, smem);
  }
  if (threadIdx.x < 32) {          // every thread holds the totals; warp 0 combines them across GPUs
    peer_allreduce<N>(rb.peer, tot);
    if (threadIdx.x == 0) {
      fin(tot);
      *rb.ticket = 0u;
      __threadfence();
    }
  }
}

// Split reductions (the big single-RHS Dslash grids).  Drawing a ticket costs every CTA an L2 atomic round trip before it
// may leave -- it has to learn whether it is the last one -- and with two 255-register CTAs per SM that ~1 us tail, during
// which the CTA's slot issues no loads, is 7-8 % of a 13 us CTA (measured on B200, profiles/r02_reduce_tail_ab.json:
// accumulate only +2.5 %, + block sums and partial stores +4-5 %, + ticket +7.6 %; a release atomic instead of
// __threadfence + atomic changes nothing).  With rb.split the CTAs store their partial and leave; this one-CTA kernel,
// launched behind the step, sums the partials in the same two-level fixed order (groups of RED_GROUP by one warp each,
// then the group sums), combines across GPUs and runs the finaliser: ~5 us instead of ~150 us of tails.
template <int N, int BLOCK, typename Fin>
__device__ __forceinline__ void reduce_finish(const ReduceBuf& rb, Fin fin) {
  __shared__ double smem[BLOCK / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = BLOCK / 32;
  const int ngrp = rb.ngroups();
  for (int g0 = w; g0 < ngrp; g0 += 4 * nw) {            // four groups per warp in flight
    double x[N][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int grp = g0 + j * nw;
      const int gsize = grp < ngrp ? min(RED_GROUP, rb.total_blocks - grp * RED_GROUP) : 0;
#pragma unroll
      for (int k = 0; k < N; ++k) {
        const double* p = rb.partial + (size_t)k * rb.total_blocks + (size_t)grp * RED_GROUP;
        x[k][j] = (lane < gsize ? __ldcg(p + lane) : 0.0) + (lane + 32 < gsize ? __ldcg(p + lane + 32) : 0.0);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int grp = g0 + j * nw;
#pragma unroll
      for (int k = 0; k < N; ++k) {
        double v = x[k][j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && grp < ngrp) rb.gpartial[(size_t)k * ngrp + grp] = v;
      }
    }
  }
  __syncthreads();                                        // one CTA: the group sums are visible to all its threads
  double tot[N];
#pragma unroll
  for (int k = 0; k < N; ++k) {
    double acc = 0.0;                                     // plain loads: written by this CTA
    for (int b = threadIdx.x; b < ngrp; b += BLOCK) acc += rb.gpartial[(size_t)k * ngrp + b];
    tot[k] = block_sum<BLOCK>(acc, smem);
  }
  if (threadIdx.x < 32) {
    peer_allreduce<N>(rb.peer, tot);
    if (threadIdx.x == 0) fin(tot);
  }
}

// Split step of a batched kernel: the warp's partial sums go to slot (rhs, site block) and nothing else happens
// (dslash_mrhs_finish_kernel sums them).  `rb` is the kernel argument itself, NOT a for_rhs() view: building that view
// copies the mailbox pointer array, which peer_allreduce indexes by lane, so the copy lives in local memory -- 36 STL.64
// per thread on the way to a reduction only one warp in 166 000 finishes (ncu, round 2: 288 B of local stores per thread =
// 3 - 4 GB of extra DRAM writes per batched reducing kernel at 48^3x48).
template <int N>
__device__ __forceinline__ void warp_partial_store(const double v[N], const ReduceBuf& rb, int rhs, int site_block) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    double x = v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) rb.partial[((size_t)rhs * N + k) * rb.total_blocks + rb.block_offset + site_block] = x;
  }
}

// Warp-synchronous variant for the multi-RHS Dslash kernels, where one WARP (32 consecutive sites of one right-hand
// side) is the reduction unit: shuffle tree -> one partial per (rhs, site block) -> two-level tickets as above (the
// warp that draws the last ticket of its group sums the group, the last group finisher sums the group sums, lane-strided
// and then the same shuffle tree: a fixed order) and lane 0 runs the finaliser.  rb must already be the for_rhs() view.
// No __syncthreads: warps of one CTA belong to different right-hand sides and may have returned early (converged).
template <int N, typename Fin>
__device__ __forceinline__ void warp_grid_reduce(double v[N], const ReduceBuf& rb, int site_block, Fin fin) {
  const int lane = threadIdx.x & 31;
  const int blk = rb.block_offset + site_block;
  const bool flat = rb.total_blocks <= RED_FLAT_MAX;
  const int grp = blk / RED_GROUP, ngrp = rb.ngroups();
  const int gsize = min(RED_GROUP, rb.total_blocks - grp * RED_GROUP);
  double s[N];
#pragma unroll
  for (int k = 0; k < N; ++k) {
    double x = v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    s[k] = x;
  }
  int role = 0;
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < N; ++k) rb.partial[(size_t)k * rb.total_blocks + blk] = s[k];
  }
  // split step (uniform over the grid): no ticket -- the release atomic behind the stores of the epilogue holds every
  // warp for an L2 round trip (batched CG, 48^3x96 x 12: EPI_M_NORM 16.9 ms against 12.5 ms for the plain EPI_M) -- a
  // small kernel behind the step sums the partials of every right-hand side (reduce_finish)
  if (rb.split) return;
  if (lane == 0) {
    if (flat) role = (draw_ticket(rb.ticket) == (unsigned int)rb.total_blocks - 1u) ? 2 : 0;
    else role = (draw_ticket(rb.gticket + grp) == (unsigned int)gsize - 1u) ? 1 : 0;
  }
  role = __shfl_sync(0xffffffffu, role, 0);
  if (role == 0) return;
  if (!flat) {
    __threadfence();
    double gs[N];
#pragma unroll
    for (int k = 0; k < N; ++k) gs[k] = group_sum(rb.partial + (size_t)k * rb.total_blocks + grp * RED_GROUP, gsize, lane);
    role = 0;
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < N; ++k) rb.gpartial[(size_t)k * ngrp + grp] = gs[k];
      rb.gticket[grp] = 0u;
      role = (draw_ticket(rb.ticket) == (unsigned int)ngrp - 1u) ? 2 : 0;
    }
    role = __shfl_sync(0xffffffffu, role, 0);
    if (role == 0) return;
  }
  __threadfence();
  double tot[N];
#pragma unroll
  for (int k = 0; k < N; ++k) {
    double acc = flat ? strided_sum(rb.partial + (size_t)k * rb.total_blocks, rb.total_blocks, lane, 32)
                      : strided_sum(rb.gpartial + (size_t)k * ngrp, ngrp, lane, 32);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    tot[k] = acc;
  }
  peer_allreduce<N>(rb.peer, tot);       // the whole warp (the butterfly left the totals in every lane)
  if (lane == 0) {
    fin(tot);
    *rb.ticket = 0u;
    __threadfence();
  }
}

}  // namespace b200
