// engine_impl.cuh -- Engine<R>: templated on the device precision (double / float).
#pragma once
#include <algorithm>
#include <cstdlib>
#include <cstdarg>
#include <thread>
#include <atomic>
#include <stdint.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "engine.cuh"
#include "dslash.cuh"
#include "blas.cuh"
#include "pack.cuh"
#include "clover_setup.cuh"
#include "halo.cuh"
#include "multishift.cuh"

namespace b200 {

#ifndef B200_DSLASH_BLOCK
#define B200_DSLASH_BLOCK 128
#endif
#ifndef B200_DSLASH_BLOCK_F
#define B200_DSLASH_BLOCK_F 128
#endif
constexpr int DSLASH_BLOCK_MAX = B200_DSLASH_BLOCK > B200_DSLASH_BLOCK_F ? B200_DSLASH_BLOCK : B200_DSLASH_BLOCK_F;
constexpr int ITER_BATCH = 8;         // iterations enqueued between two status polls
constexpr size_t STAGING_BYTES = 256u << 20;
constexpr size_t PIN_BYTES_DEFAULT = 32u << 20;   // one pinned bounce buffer of the pageable-host copy pipeline (two per engine)

// Copy with non-temporal stores: the destination of a download is a gigabyte of pageable user memory that is not read
// again soon, so ordinary stores would first read every destination line into the cache (read-for-ownership) and double
// the memory traffic of the copy.  Falls back to memcpy without AVX2.
#if defined(__x86_64__)
__attribute__((target("avx2"))) inline void stream_copy_avx2(void* dst, const void* src, size_t bytes) {
  char* d = (char*)dst; const char* s = (const char*)src;
  const size_t head = ((uintptr_t)d & 31) ? 32 - ((uintptr_t)d & 31) : 0;
  if (head >= bytes) { memcpy(d, s, bytes); return; }
  memcpy(d, s, head); d += head; s += head; bytes -= head;
  const size_t n32 = bytes / 32;
  for (size_t i = 0; i < n32; ++i) _mm256_stream_si256((__m256i*)d + i, _mm256_loadu_si256((const __m256i*)s + i));
  _mm_sfence();
  memcpy(d + n32 * 32, s + n32 * 32, bytes - n32 * 32);
}
inline void stream_copy(void* dst, const void* src, size_t bytes) {
  static const bool avx2 = __builtin_cpu_supports("avx2");
  if (avx2 && bytes >= 4096) stream_copy_avx2(dst, src, bytes); else memcpy(dst, src, bytes);
}
#else
inline void stream_copy(void* dst, const void* src, size_t bytes) { memcpy(dst, src, bytes); }
#endif

// memcpy with a small team of threads: one core moves ~10 GB/s, PCIe 5 x16 wants ~50.  nt: non-temporal stores (downloads).
inline void parallel_memcpy(void* dst, const void* src, size_t bytes, int nthreads, size_t serial_below, bool nt = false) {
  auto cp = [nt](void* d, const void* s, size_t n) { if (nt) stream_copy(d, s, n); else memcpy(d, s, n); };
  if (nthreads <= 1 || bytes < serial_below) { cp(dst, src, bytes); return; }
  const size_t per = ((bytes + nthreads - 1) / nthreads + 4095) & ~(size_t)4095;
  std::vector<std::thread> team;
  for (int t = 1; t < nthreads; ++t) {
    const size_t off = per * t;
    if (off >= bytes) break;
    team.emplace_back([=] { cp((char*)dst + off, (const char*)src + off, std::min(per, bytes - off)); });
  }
  cp(dst, src, std::min(per, bytes));
  for (auto& th : team) th.join();
}

// True if every byte of a host buffer is zero.  The propagator call site hands the solver psi = zero as the initial guess
// (quarkprop4_w.cc:74: "LatticeFermion psi = zero"), and scanning a gigabyte with the thread team (read-only, early exit at
// the first non-zero word -- a real initial guess costs one cache line) is 2-3x cheaper than bouncing it over PCIe just to
// hold zeros on the device.
inline bool host_all_zero(const void* p, size_t bytes, int nthreads) {
  if (bytes % 8 != 0 || ((uintptr_t)p & 7)) return false;
  const unsigned long long* w = (const unsigned long long*)p;
  const size_t n = bytes / 8;
  if (n == 0 || w[0] != 0ull || w[n - 1] != 0ull || w[n / 2] != 0ull) return false;
  std::atomic<bool> nonzero(false);
  auto scan = [&](size_t lo, size_t hi) {
    for (size_t i = lo; i < hi && !nonzero.load(std::memory_order_relaxed); i += 512) {
      unsigned long long acc = 0ull;
      const size_t e = std::min(hi, i + 512);
      for (size_t j = i; j < e; ++j) acc |= w[j];
      if (acc) nonzero.store(true, std::memory_order_relaxed);
    }
  };
  const int nt = std::max(1, nthreads);
  const size_t per = (n + nt - 1) / nt;
  std::vector<std::thread> team;
  for (int t = 1; t < nt; ++t) if (per * t < n) team.emplace_back(scan, per * t, std::min(n, per * (t + 1)));
  scan(0, std::min(n, per));
  for (auto& th : team) th.join();
  return !nonzero.load();
}

// status != nullptr: a solver launch -- return at once when the solve has stopped / when slot run_if is clear
template <typename R, int BLOCK>
__global__ void __launch_bounds__(BLOCK) clover_kernel(const Cx<R>* __restrict__ in, Cx<R>* __restrict__ out,
                                                      const Cx<R>* __restrict__ clov, int Vh, size_t fstride,
                                                      const int* __restrict__ status = nullptr, int run_if = 0) {
  if (status && (status[ST_STOP] != 0 || status[ST_BREAKDOWN] != 0)) return;
  if (status && run_if && status[run_if] == 0) return;
  const int idx = blockIdx.x * BLOCK + threadIdx.x;
  if (idx >= Vh) return;
  in += blockIdx.y * fstride; out += blockIdx.y * fstride;
#pragma unroll
  for (int b = 0; b < 2; ++b) {
    Cx<R> xi[6], o[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) xi[k] = ldg(in + (size_t)(6 * b + k) * Vh + idx);
    clover_block<R>(o, xi, clov + (size_t)(18 * b) * Vh + idx, Vh, make_l2_policy().stream);
#pragma unroll
    for (int k = 0; k < 6; ++k) out[(size_t)(6 * b + k) * Vh + idx] = o[k];
  }
}

struct ScalarSet { int n; int slots[12]; double vals[12]; int reset_status; int rhs; };
__global__ void set_scalars_kernel(double* scal, int* status, ScalarSet s);
__global__ void l2_policy_kernel(L2Policy* out);
__global__ void scale_planes_kernel(double2* p, int nplanes, size_t stride, size_t off, int count, double f);
__global__ void scale_planes_kernel(float2* p, int nplanes, size_t stride, size_t off, int count, double f);

template <typename R>
class Engine : public EngineBase {
 public:
  typedef Cx<R> C;
  static constexpr int DSLASH_BLOCK = sizeof(R) == 4 ? B200_DSLASH_BLOCK_F : B200_DSLASH_BLOCK;   // tuned per precision
  explicit Engine(const Config& c) { cfg = c; }
  ~Engine() override { destroy(); }

  // ------------------------------------------------------------------ state
  C* gauge = nullptr;       // [4][2][NG][Vh]
  int recon = 0;            // 18 / 12 once loaded
  double aniso_[4] = {1, 1, 1, 1};
  int t_boundary_ = 1;
  LinkScale ls{};
  C* clov = nullptr;        // [2][36][Vh]
  C* invclov = nullptr;     // [36][Vh]  (cb 0)
  double* tr_log = nullptr; // [Vh] log|det A_ee| per even site (make_clover only)
  bool have_trlog = false;
  // symmetric preconditioning (SymEvenOddPrecCloverLinOp, seoprec_clover_linop_w.cc:16-41): needs A_oo^-1 as well
  int sym = 0;              // 0: M = A_oo - 1/4 D A_ee^-1 D ; 1: M = 1 - 1/4 A_oo^-1 D A_ee^-1 D
  C* invclov_oo = nullptr;  // [36][Vh] (cb 1), derived on the device from clov whenever sym is on
  double* tr_log_oo = nullptr;
  // multi-shift CG (MInvCG2_a)
  MsState* ms_dev = nullptr; b200_field* ms_p = nullptr; C* ms_psi = nullptr;
  double* scal = nullptr; int* status = nullptr;
  double* partial = nullptr; unsigned int* ticket = nullptr; size_t partial_cap = 0;   // partial: doubles; ticket: [MAX_RHS]
  double* gpartial = nullptr; unsigned int* gticket = nullptr;                          // two-level reductions (reduce.cuh)
  double* h_scal = nullptr; int* h_status = nullptr;   // pinned; [MAX_RHS][S_COUNT] and 3 slots of [MAX_RHS][ST_COUNT] (2 for the solver poll, 1 for comm_check)
  cudaEvent_t ev_poll[2] = {nullptr, nullptr}, ev_t0 = nullptr, ev_t1 = nullptr;
  void* staging = nullptr;
  // Host buffers the caller hands over (QDP++ fields) are pageable: a plain cudaMemcpy moves them at ~10 GB/s through the
  // driver's own bounce buffer.  We bounce them ourselves: a team of host threads fills one pinned buffer while the DMA
  // engine drains the other (and the reverse for downloads).  Pinned / registered host memory takes the direct path.
  void* pin[2] = {nullptr, nullptr};
  cudaEvent_t pin_ev[2] = {nullptr, nullptr};
  int pin_next = 0;
  int copy_threads = 4;
  bool nt_copy = true;      // downloads leave the pinned bounce buffer with non-temporal stores (B200_NT_COPY=0: plain memcpy)
  size_t PIN_BYTES = PIN_BYTES_DEFAULT;   // B200_PIN_KB shrinks it (tests: many pieces and the thread team on small lattices)
  L2Policy l2pol{};         // createpolicy descriptors (evict_last for neighbour spinors, evict_first for streams), made once
  b200_field* ws[9] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  Halo<R> halo;
  int blas_grid = 148 * 8;
  int num_sms = 148;
  bool owns_stream = true, owns_scalars = true;   // false once a mixed-precision partner lent us its stream / scalar block
  long long operator_epoch = 0;                   // bumped whenever gauge or clover change (the fp32 twin re-syncs on it)
  // fixed-iteration (benchmark) state
  C* it_psi = nullptr; const C* it_chi = nullptr; int it_k = 0, it_nb = 1;

  size_t nelem() const { return (size_t)12 * g.Vh; }
  bool split() const { return g.tsplit || g.zsplit; }
  double nranks() const { return (double)cfg.pgrid[2] * cfg.pgrid[3]; }

  // ------------------------------------------------------------------ lifecycle
  int init() override {
    for (int i = 0; i < 4; ++i) {
      if (cfg.pgrid[i] < 1 || cfg.gdims[i] % cfg.pgrid[i] != 0) { set_error("global dim %d not divisible by grid", i); return B200_ERR_ARG; }
      cfg.ldims[i] = cfg.gdims[i] / cfg.pgrid[i];
      // checkerboarding needs even GLOBAL dims (shift_table_scalar.cc:23-28); we also need even LOCAL dims so
      // that local parity == global parity on every rank (cf. cpp_dslash_parscalar_64bit.cc:187-212)
      if (cfg.ldims[i] % 2 != 0 || cfg.ldims[i] < 2) { set_error("local lattice extent %d (dim %d) must be even and >= 2", cfg.ldims[i], i); return B200_ERR_ARG; }
    }
    if (cfg.pgrid[0] != 1 || cfg.pgrid[1] != 1) { set_error("the process grid may split T and Z only (1 x 1 x Pz x Pt)"); return B200_ERR_ARG; }
    if (cfg.pgrid[2] * cfg.pgrid[3] > 8) { set_error("at most 8 ranks (one NVSwitch node)"); return B200_ERR_ARG; }
    if (cfg.pgrid[2] * cfg.pgrid[3] > 1 && !cfg.have_comm) { set_error("a b200_comm is required for a split lattice"); return B200_ERR_COMM; }
    // the kernels' division-free site decode (FastDiv, common.cuh) is exact for site counts below 2^28 per checkerboard and rank
    if ((long long)(cfg.ldims[0] / 2) * cfg.ldims[1] * cfg.ldims[2] * cfg.ldims[3] >= (1ll << 28)) {
      set_error("local lattice %d x %d x %d x %d has 2^28 or more sites per checkerboard: split it over more ranks", cfg.ldims[0], cfg.ldims[1], cfg.ldims[2], cfg.ldims[3]);
      return B200_ERR_ARG;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device: this engine has no CPU fallback"); return B200_ERR_CUDA; }
    B200_CUDA(cudaSetDevice(cfg.device));
    cudaDeviceProp prop;
    B200_CUDA(cudaGetDeviceProperties(&prop, cfg.device));
    if (prop.major < 10) { set_error("device %s is sm_%d%d; this library is built for sm_100a only", prop.name, prop.major, prop.minor); return B200_ERR_CUDA; }
    blas_grid = prop.multiProcessorCount * 8;
    num_sms = prop.multiProcessorCount;
    if (const char* e = getenv("B200_MRHS_L2_KB")) l2_budget = atol(e) << 10;
    if (const char* e = getenv("B200_SPLIT_MIN_BLOCKS")) split_min_blocks = atoi(e);
    if (const char* e = getenv("B200_MRHS_YCHUNK")) y_chunks = atoi(e);
    g.Lxh = cfg.ldims[0] / 2; g.Ly = cfg.ldims[1]; g.Lz = cfg.ldims[2]; g.Lt = cfg.ldims[3];
    g.S3h = g.Lxh * g.Ly * g.Lz;
    g.Vh = g.S3h * g.Lt;
    g.tsplit = cfg.pgrid[3] > 1 ? 1 : 0;
    g.zsplit = cfg.pgrid[2] > 1 ? 1 : 0;
    g.SZh = g.Lxh * g.Ly * g.Lt;
    B200_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    B200_CUDA(cudaMalloc(&scal, sizeof(double) * S_COUNT * MAX_RHS));
    B200_CUDA(cudaMalloc(&status, sizeof(int) * ST_COUNT * MAX_RHS));
    B200_CUDA(cudaMemsetAsync(scal, 0, sizeof(double) * S_COUNT * MAX_RHS, stream));
    B200_CUDA(cudaMemsetAsync(status, 0, sizeof(int) * ST_COUNT * MAX_RHS, stream));
    { int rc = ensure_partial(4 * (size_t)std::max((g.Vh + DSLASH_BLOCK - 1) / DSLASH_BLOCK + 8, blas_grid)); if (rc) return rc; }
    B200_CUDA(cudaMalloc(&ticket, sizeof(unsigned int) * MAX_RHS));
    B200_CUDA(cudaMemsetAsync(ticket, 0, sizeof(unsigned int) * MAX_RHS, stream));
    B200_CUDA(cudaHostAlloc(&h_scal, sizeof(double) * S_COUNT * MAX_RHS, cudaHostAllocDefault));
    B200_CUDA(cudaHostAlloc(&h_status, sizeof(int) * ST_COUNT * MAX_RHS * 3, cudaHostAllocDefault));
    for (int i = 0; i < 2; ++i) B200_CUDA(cudaEventCreateWithFlags(&ev_poll[i], cudaEventDisableTiming));
    B200_CUDA(cudaEventCreate(&ev_t0));
    B200_CUDA(cudaEventCreate(&ev_t1));
    B200_CUDA(cudaMalloc(&staging, STAGING_BYTES));
    l2_policy_kernel<<<1, 1, 0, stream>>>((L2Policy*)staging);
    B200_CUDA(cudaMemcpyAsync(&l2pol, staging, sizeof(L2Policy), cudaMemcpyDeviceToHost, stream));
    B200_CUDA(cudaStreamSynchronize(stream));
    if (const char* e = getenv("B200_PIN_KB")) PIN_BYTES = std::max<size_t>(4096, (size_t)atol(e) << 10);
    for (int i = 0; i < 2; ++i) {
      B200_CUDA(cudaHostAlloc(&pin[i], PIN_BYTES, cudaHostAllocDefault));
      B200_CUDA(cudaEventCreateWithFlags(&pin_ev[i], cudaEventDisableTiming));
    }
    // the ranks of one box share its cores: all of them, split over the ranks, at most 16 per rank (the host has nothing else
    // to do while a field is copied).  16-core box, 1 rank: gauge upload 189 / 161 / 140 ms with 8 / 12 / 16 threads, 1 GB
    // fermion down 33 / 29 / 28 ms (profiles/r02_copy_threads.json)
    copy_threads = (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency() / (unsigned)(cfg.pgrid[2] * cfg.pgrid[3])));
    if (const char* e = getenv("B200_COPY_THREADS")) copy_threads = std::max(0, atoi(e));
    if (const char* e = getenv("B200_NT_COPY")) nt_copy = atoi(e) != 0;
    host_copy_threads = copy_threads;   // 0: plain cudaMemcpy from pageable memory
    if (split()) { int rc = halo.init(cfg, g, stream); if (rc) return rc; halo.status_dev = status; }
    B200_CUDA(cudaStreamSynchronize(stream));
    return B200_OK;
  }

  void destroy() {
    if (!stream) return;
    cudaSetDevice(cfg.device);
    cudaStreamSynchronize(stream);
    halo.destroy();
    for (auto& f : ws) if (f) { field_free(f); f = nullptr; }
    if (ms_p) { field_free(ms_p); ms_p = nullptr; }
    cudaFree(gauge); cudaFree(clov); cudaFree(invclov); cudaFree(tr_log); cudaFree(invclov_oo); cudaFree(tr_log_oo); cudaFree(ms_dev);
    if (owns_scalars) { cudaFree(scal); cudaFree(status); }
    cudaFree(partial); cudaFree(gpartial); cudaFree(gticket); cudaFree(ticket); cudaFree(staging);
    for (int i = 0; i < 2; ++i) { if (pin[i]) cudaFreeHost(pin[i]); if (pin_ev[i]) cudaEventDestroy(pin_ev[i]); pin[i] = nullptr; pin_ev[i] = nullptr; }
    cudaFreeHost(h_scal); cudaFreeHost(h_status);
    for (int i = 0; i < 2; ++i) if (ev_poll[i]) cudaEventDestroy(ev_poll[i]);
    for (auto& e : trace_ev) cudaEventDestroy(e);
    trace_ev.clear();
    if (ev_t0) cudaEventDestroy(ev_t0);
    if (ev_t1) cudaEventDestroy(ev_t1);
    if (owns_stream) cudaStreamDestroy(stream);
    stream = nullptr;
  }

  // reduction scratch: `doubles` partial sums (grown on demand by the batched kernels: N x nrhs x blocks), plus the
  // group sums and group tickets of the two-level reductions (reduce.cuh): one per RED_GROUP partials
  int ensure_partial(size_t doubles) {
    if (doubles <= partial_cap) return B200_OK;
    if (partial) { B200_CUDA(cudaStreamSynchronize(stream)); cudaFree(partial); cudaFree(gpartial); cudaFree(gticket); partial = nullptr; gpartial = nullptr; gticket = nullptr; }
    const size_t ng = doubles / RED_GROUP + 8 * MAX_RHS;
    B200_CUDA(cudaMalloc(&partial, sizeof(double) * doubles));
    B200_CUDA(cudaMalloc(&gpartial, sizeof(double) * ng));
    B200_CUDA(cudaMalloc(&gticket, sizeof(unsigned int) * ng));
    B200_CUDA(cudaMemsetAsync(gticket, 0, sizeof(unsigned int) * ng, stream));
    partial_cap = doubles;
    return B200_OK;
  }

  int sync() override { B200_CUDA(cudaSetDevice(cfg.device)); B200_CUDA(cudaStreamSynchronize(stream)); return comm_check(); }

  // Split lattices: a peer wait that ran out of its spin budget (status 90 = halo flag, 91 = reduction mailbox) left stale
  // ghost / mailbox data behind.  The solver loops see that in their status poll; every other path calls this before a
  // result reaches the host (fetch_scalars, the end of a download, b200_sync), so a lost peer is B200_ERR_COMM, never a
  // silently wrong number.
  int comm_check() {
    if (!split()) return B200_OK;
    int* hs3 = h_status + 2 * ST_COUNT * MAX_RHS;
    B200_CUDA(cudaMemcpyAsync(hs3, status, sizeof(int) * ST_COUNT * MAX_RHS, cudaMemcpyDeviceToHost, stream));
    B200_CUDA(cudaStreamSynchronize(stream));
    for (int r = 0; r < MAX_RHS; ++r) if (hs3[r * ST_COUNT + ST_BREAKDOWN] >= 90) return comm_timeout(hs3[r * ST_COUNT + ST_BREAKDOWN]);
    return B200_OK;
  }

  // b200_dev_time_solver_kernels: while trace_n >= 0 every hot-loop launch is followed by an event on the stream
  std::vector<cudaEvent_t> trace_ev; int trace_n = -1;
  void mark() { if (trace_n >= 0 && trace_n < (int)trace_ev.size()) cudaEventRecord(trace_ev[trace_n++], stream); }

  int launched(const char* what) {
    ++launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("launch of %s failed: %s", what, cudaGetErrorString(e)); return B200_ERR_CUDA; }
    return B200_OK;
  }

  ReduceBuf make_red(int block_offset, int total_blocks) {
    ReduceBuf rb;
    rb.partial = partial; rb.ticket = ticket; rb.gpartial = gpartial; rb.gticket = gticket; rb.block_offset = block_offset; rb.total_blocks = total_blocks;
    rb.split = 0;
    rb.peer = halo.peer_reduce();
    return rb;
  }

  // ------------------------------------------------------------------ host <-> device copies
  static bool is_pageable(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return at.type == cudaMemoryTypeUnregistered;
  }
  // Enqueue host -> device; on return the host buffer has been read completely (pageable) or the copy is in the stream (pinned).
  int h2d(void* dst_dev, const void* src, size_t bytes) {
    if (copy_threads == 0 || !is_pageable(src)) {
      B200_CUDA(cudaMemcpyAsync(dst_dev, src, bytes, cudaMemcpyHostToDevice, stream));
      return B200_OK;
    }
    for (size_t off = 0; off < bytes; off += PIN_BYTES) {
      const size_t n = std::min(PIN_BYTES, bytes - off);
      const int k = pin_next; pin_next ^= 1;
      B200_CUDA(cudaEventSynchronize(pin_ev[k]));                 // the DMA that last read this bounce buffer is done
      parallel_memcpy(pin[k], (const char*)src + off, n, copy_threads, PIN_BYTES / 8);
      B200_CUDA(cudaMemcpyAsync((char*)dst_dev + off, pin[k], n, cudaMemcpyHostToDevice, stream));
      B200_CUDA(cudaEventRecord(pin_ev[k], stream));
    }
    return B200_OK;
  }
  // Device -> host, complete on return.
  int d2h(void* dst, const void* src_dev, size_t bytes) {
    if (copy_threads == 0 || !is_pageable(dst)) {
      B200_CUDA(cudaMemcpyAsync(dst, src_dev, bytes, cudaMemcpyDeviceToHost, stream));
      B200_CUDA(cudaStreamSynchronize(stream));
      return B200_OK;
    }
    size_t prev_off = 0, prev_n = 0; int prev_k = -1;
    for (size_t off = 0; off < bytes || prev_k >= 0; off += PIN_BYTES) {
      int k = -1; size_t n = 0;
      if (off < bytes) {
        n = std::min(PIN_BYTES, bytes - off);
        k = pin_next; pin_next ^= 1;
        B200_CUDA(cudaEventSynchronize(pin_ev[k]));
        B200_CUDA(cudaMemcpyAsync(pin[k], (const char*)src_dev + off, n, cudaMemcpyDeviceToHost, stream));
        B200_CUDA(cudaEventRecord(pin_ev[k], stream));
      }
      if (prev_k >= 0) {                                          // drain the previous piece while this one is in flight
        B200_CUDA(cudaEventSynchronize(pin_ev[prev_k]));
        parallel_memcpy((char*)dst + prev_off, pin[prev_k], prev_n, copy_threads, PIN_BYTES / 8, nt_copy);
      }
      prev_k = k; prev_off = off; prev_n = n;
    }
    return B200_OK;
  }

  // ------------------------------------------------------------------ host <-> device reordering
  template <typename H, int NR, int NPL, typename Map>
  int upload_aos(const H* src, int nsites, C* dst, size_t stride, Map map, double scale) {
    const size_t rec = (size_t)NR * sizeof(H);
    int chunk = (int)std::min<size_t>((size_t)nsites, (STAGING_BYTES / rec) / PACK_SITES * PACK_SITES);
    for (int off = 0; off < nsites; off += chunk) {
      const int n = std::min(chunk, nsites - off);
      { int rch = h2d(staging, (const char*)src + (size_t)off * rec, (size_t)n * rec); if (rch) return rch; }
      aos_to_soa_kernel<H, R, NR, NPL, Map><<<(n + PACK_SITES - 1) / PACK_SITES, PACK_BLOCK, 0, stream>>>(
          (const H*)staging, dst, n, stride, (size_t)off, map, scale);
      int rc = launched("aos_to_soa"); if (rc) return rc;
      if (NR == 18 && r12_check) {   // 12-real compression: every link must be SU(3) up to the one phase the engine carries itself
        recon12_check_kernel<H><<<(n + 255) / 256, 256, 0, stream>>>((const H*)staging, n, (size_t)off, r12_flip_lo, r12_flip_hi, r12_check);
        rc = launched("recon12_check"); if (rc) return rc;
      }
      // the staging buffer is reused by the next chunk: same stream, so ordering is implicit
    }
    return B200_OK;
  }
  template <typename H, int NR, int NPL, typename Map>
  int download_aos(H* dst, int nsites, const C* src, size_t stride, Map map) {
    const size_t rec = (size_t)NR * sizeof(H);
    int chunk = (int)std::min<size_t>((size_t)nsites, (STAGING_BYTES / rec) / PACK_SITES * PACK_SITES);
    for (int off = 0; off < nsites; off += chunk) {
      const int n = std::min(chunk, nsites - off);
      soa_to_aos_kernel<H, R, NR, NPL, Map><<<(n + PACK_SITES - 1) / PACK_SITES, PACK_BLOCK, 0, stream>>>(
          (H*)staging, src, n, stride, (size_t)off, map);
      int rc = launched("soa_to_aos"); if (rc) return rc;
      // the next chunk's kernel overwrites the device staging buffer: d2h returns only when this chunk has left it
      { int rch = d2h((char*)dst + (size_t)off * rec, staging, (size_t)n * rec); if (rch) return rch; }
    }
    B200_CUDA(cudaStreamSynchronize(stream));
    return comm_check();
  }

  unsigned long long* r12_check = nullptr;    // device word: max |reconstructed row 2 - given row 2| of the upload in progress
  size_t r12_flip_lo = 0, r12_flip_hi = 0;
  int load_gauge(const void* const u[4], int host_prec, const double aniso[4], int t_boundary, int recon_) override {
    B200_CUDA(cudaSetDevice(cfg.device));
    if (recon_ != B200_RECONS_NONE && recon_ != B200_RECONS_12) { set_error("reconstruct must be 18 or 12"); return B200_ERR_ARG; }
    if (host_prec != B200_SINGLE && host_prec != B200_DOUBLE) { set_error("host_prec must be 4 or 8"); return B200_ERR_ARG; }
    if (t_boundary != 1 && t_boundary != -1) { set_error("t_boundary must be +1 or -1"); return B200_ERR_ARG; }
    for (int mu = 0; mu < 4; ++mu) if (!u[mu]) { set_error("null gauge pointer"); return B200_ERR_ARG; }
    const int NG = recon_ / 2;
    if (gauge) { cudaFree(gauge); gauge = nullptr; }
    B200_CUDA(cudaMalloc(&gauge, sizeof(C) * 8 * (size_t)NG * g.Vh));
    recon = recon_; t_boundary_ = t_boundary;
    const bool last_rank = cfg.pcoord[3] == cfg.pgrid[3] - 1;
    unsigned long long* chk = nullptr;
    if (recon == 12) { B200_CUDA(cudaMalloc(&chk, sizeof(unsigned long long))); B200_CUDA(cudaMemsetAsync(chk, 0, sizeof(unsigned long long), stream)); }
    for (int mu = 0; mu < 4; ++mu) {
      aniso_[mu] = aniso[mu];
      ls.aniso[mu] = aniso[mu];
      for (int par = 0; par < 2; ++par) {
        r12_check = chk;
        const bool flip = mu == 3 && t_boundary == -1 && last_rank;      // the -1 of the antiperiodic T boundary lives on the last slice
        r12_flip_lo = flip ? (size_t)(g.Lt - 1) * g.S3h : 0; r12_flip_hi = flip ? (size_t)g.Vh : 0;
        C* dst = gauge + (size_t)(mu * 2 + par) * NG * g.Vh;
        int rc;
        const double sc = (recon == 18) ? aniso[mu] : 1.0;
        if (host_prec == B200_DOUBLE) {
          const double* src = (const double*)u[mu] + (size_t)par * g.Vh * 18;
          rc = (recon == 18) ? upload_aos<double, 18, 9>(src, g.Vh, dst, (size_t)g.Vh, MapIdentity(), sc)
                             : upload_aos<double, 18, 6>(src, g.Vh, dst, (size_t)g.Vh, MapIdentity(), sc);
        } else {
          const float* src = (const float*)u[mu] + (size_t)par * g.Vh * 18;
          rc = (recon == 18) ? upload_aos<float, 18, 9>(src, g.Vh, dst, (size_t)g.Vh, MapIdentity(), sc)
                             : upload_aos<float, 18, 6>(src, g.Vh, dst, (size_t)g.Vh, MapIdentity(), sc);
        }
        r12_check = nullptr;
        if (rc) { cudaFree(chk); return rc; }
        if (recon == 12 && mu == 3 && t_boundary == -1 && last_rank) {
          // strip the antiperiodic phase so that rows 0,1 are rows of an SU(3) matrix again; the kernel
          // re-applies it (LinkScale::bc_t), as the QUDA / QPhiX adapters do
          // (syssolver_linop_clover_quda_w.h:147-156, syssolver_linop_clover_qphix_w.h:107-116)
          scale_planes_kernel<<<(g.S3h + 255) / 256, 256, 0, stream>>>(dst, 6, (size_t)g.Vh, (size_t)(g.Lt - 1) * g.S3h, g.S3h, -1.0);
          rc = launched("scale_planes"); if (rc) return rc;
        }
      }
    }
    ls.bc_t = t_boundary;
    ls.t_is_last = last_rank ? 1 : 0;
    ++operator_epoch;
    B200_CUDA(cudaStreamSynchronize(stream));
    if (chk) {
      unsigned long long bits = 0;
      B200_CUDA(cudaMemcpy(&bits, chk, sizeof(bits), cudaMemcpyDeviceToHost));
      cudaFree(chk);
      double dev; memcpy(&dev, &bits, sizeof(dev));
      const double tol = host_prec == B200_DOUBLE ? 1e-9 : 1e-3;
      if (!(dev <= tol)) {
        cudaFree(gauge); gauge = nullptr;
        set_error("RECONS_12 needs SU(3) links whose only extra phase is the antiperiodic T boundary (t_boundary = %d): the third row "
                  "rebuilt from rows 0,1 differs from the given one by %.3e -- spatial fermion boundary phases, a wrong AntiPeriodicT "
                  "or non-unitary links; use RECONS_NONE", t_boundary, dev);
        return B200_ERR_ARG;
      }
    }
    return B200_OK;
  }

  int load_clover(const void* clov_h, const void* invclov_h, int host_prec) override {
    B200_CUDA(cudaSetDevice(cfg.device));
    if (!clov_h || !invclov_h) { set_error("null clover pointer"); return B200_ERR_ARG; }
    if (host_prec != B200_SINGLE && host_prec != B200_DOUBLE) { set_error("host_prec must be 4 or 8"); return B200_ERR_ARG; }
    int rc = alloc_clover(); if (rc) return rc;
    for (int par = 0; par < 2; ++par) {
      C* dst = clov + (size_t)par * 36 * g.Vh;
      if (host_prec == B200_DOUBLE) rc = upload_aos<double, 72, 36>((const double*)clov_h + (size_t)par * g.Vh * 72, g.Vh, dst, (size_t)g.Vh, MapClover(), 1.0);
      else rc = upload_aos<float, 72, 36>((const float*)clov_h + (size_t)par * g.Vh * 72, g.Vh, dst, (size_t)g.Vh, MapClover(), 1.0);
      if (rc) return rc;
    }
    if (host_prec == B200_DOUBLE) rc = upload_aos<double, 72, 36>((const double*)invclov_h, g.Vh, invclov, (size_t)g.Vh, MapClover(), 1.0);
    else rc = upload_aos<float, 72, 36>((const float*)invclov_h, g.Vh, invclov, (size_t)g.Vh, MapClover(), 1.0);
    if (rc) return rc;
    have_trlog = false;
    rc = refresh_sym(); if (rc) return rc;
    ++operator_epoch;
    B200_CUDA(cudaStreamSynchronize(stream));
    return B200_OK;
  }

  // A_oo^-1 (and log|det A_oo|) from the odd half of clov: invclov.choles(1), seoprec_clover_linop_w.cc:31-33
  int refresh_sym() {
    if (!sym || !clov) return B200_OK;
    if (!invclov_oo) B200_CUDA(cudaMalloc(&invclov_oo, sizeof(C) * 36 * (size_t)g.Vh));
    if (!tr_log_oo) B200_CUDA(cudaMalloc(&tr_log_oo, sizeof(double) * (size_t)g.Vh));
    B200_CUDA(cudaMemcpyAsync(invclov_oo, clov + (size_t)36 * g.Vh, sizeof(C) * 36 * (size_t)g.Vh, cudaMemcpyDeviceToDevice, stream));
    ldagdlinv_kernel<R><<<(2 * g.Vh + CLOV_BLOCK - 1) / CLOV_BLOCK, CLOV_BLOCK, 0, stream>>>(invclov_oo, tr_log_oo, g.Vh);
    return launched("ldagdlinv(oo)");
  }
  int set_preconditioning(int mode) override {
    B200_CUDA(cudaSetDevice(cfg.device));
    if (mode != B200_PRECOND_ASYMMETRIC && mode != B200_PRECOND_SYMMETRIC) { set_error("preconditioning must be B200_PRECOND_ASYMMETRIC or B200_PRECOND_SYMMETRIC"); return B200_ERR_ARG; }
    if (mode == sym) return B200_OK;
    sym = mode;
    int rc = refresh_sym(); if (rc) return rc;
    ++operator_epoch;
    it_psi = nullptr;
    B200_CUDA(cudaStreamSynchronize(stream));
    return B200_OK;
  }

  // Twisted-mass term of the clover operators: chi += (+/-) mu i gamma_5 psi for PLUS / MINUS
  // (eoprec_clover_linop_w.cc:174-184, seoprec_clover_linop_w.cc:174-184; CloverFermActParams::twisted_m).  0 = none.
  double twisted_m = 0.0;
  int set_twisted_mass(double mu) override {
    if (!(mu == mu)) { set_error("b200_set_twisted_mass: NaN"); return B200_ERR_ARG; }
    if (mu != twisted_m) { twisted_m = mu; ++operator_epoch; it_psi = nullptr; }
    return B200_OK;
  }

  int alloc_clover() {
    if (!clov) B200_CUDA(cudaMalloc(&clov, sizeof(C) * 72 * (size_t)g.Vh));
    if (!invclov) B200_CUDA(cudaMalloc(&invclov, sizeof(C) * 36 * (size_t)g.Vh));
    if (!tr_log) B200_CUDA(cudaMalloc(&tr_log, sizeof(double) * (size_t)g.Vh));
    return B200_OK;
  }

  int make_clover(double diag_mass, double cr, double ct, int aniso, int t_dir) override {
    B200_CUDA(cudaSetDevice(cfg.device));
    if (!gauge) { set_error("b200_make_clover: load the gauge field first"); return B200_ERR_STATE; }
    if (t_dir < 0 || t_dir > 3) { set_error("t_dir out of range"); return B200_ERR_ARG; }
    int rc = alloc_clover(); if (rc) return rc;
    CloverSetupArgs<R> a;
    a.gauge = gauge; a.recon12 = (recon == 12); a.g = g;
    for (int mu = 0; mu < 4; ++mu) a.inv_aniso[mu] = (recon == 18) ? 1.0 / aniso_[mu] : 1.0;
    a.bc_t = (recon == 12) ? t_boundary_ : 1; a.t_is_last = ls.t_is_last;
    a.diag_mass = diag_mass; a.cr = cr; a.ct = ct; a.aniso = aniso; a.t_dir = t_dir;
    a.ghost_links = split() ? halo.gauge_ghost() : nullptr;
    a.ghost_links_z = split() ? halo.gauge_ghost_z() : nullptr;
    a.parity = 0; a.clov_out = nullptr;
    if (split()) { rc = halo.exchange_gauge_ghost(a, launches); if (rc) return rc; }
    for (int par = 0; par < 2; ++par) {
      a.parity = par; a.clov_out = clov + (size_t)par * 36 * g.Vh;
      make_clover_kernel<R><<<(g.Vh + CLOV_SITES - 1) / CLOV_SITES, dim3(CLOV_SITES, 6), 0, stream>>>(a);
      rc = launched("make_clover"); if (rc) return rc;
    }
    B200_CUDA(cudaMemcpyAsync(invclov, clov, sizeof(C) * 36 * (size_t)g.Vh, cudaMemcpyDeviceToDevice, stream));
    ldagdlinv_kernel<R><<<(2 * g.Vh + CLOV_BLOCK - 1) / CLOV_BLOCK, CLOV_BLOCK, 0, stream>>>(invclov, tr_log, g.Vh);
    rc = launched("ldagdlinv"); if (rc) return rc;
    have_trlog = true;
    rc = refresh_sym(); if (rc) return rc;
    ++operator_epoch;
    B200_CUDA(cudaStreamSynchronize(stream));
    return B200_OK;
  }

  int get_clover(void* clov_h, void* invclov_h, int host_prec) override {
    B200_CUDA(cudaSetDevice(cfg.device));
    if (!clov) { set_error("no clover term loaded"); return B200_ERR_STATE; }
    int rc = B200_OK;
    if (clov_h) for (int par = 0; par < 2 && !rc; ++par) {
      const C* src = clov + (size_t)par * 36 * g.Vh;
      if (host_prec == B200_DOUBLE) rc = download_aos<double, 72, 36>((double*)clov_h + (size_t)par * g.Vh * 72, g.Vh, src, (size_t)g.Vh, MapClover());
      else rc = download_aos<float, 72, 36>((float*)clov_h + (size_t)par * g.Vh * 72, g.Vh, src, (size_t)g.Vh, MapClover());
    }
    if (invclov_h && !rc) {
      if (host_prec == B200_DOUBLE) rc = download_aos<double, 72, 36>((double*)invclov_h, g.Vh, invclov, (size_t)g.Vh, MapClover());
      else rc = download_aos<float, 72, 36>((float*)invclov_h, g.Vh, invclov, (size_t)g.Vh, MapClover());
    }
    return rc;
  }

  int clover_logdet(double* out, int cb) override {
    B200_CUDA(cudaSetDevice(cfg.device));
    if (cb != 0 && cb != 1) { set_error("cb must be 0 or 1"); return B200_ERR_ARG; }
    if (cb == 0 && !have_trlog) { set_error("tr log is only available after b200_make_clover"); return B200_ERR_STATE; }
    if (cb == 1 && (!sym || !tr_log_oo)) { set_error("log det A_oo needs symmetric preconditioning and a clover term"); return B200_ERR_STATE; }
    { int rcb = set_batch(1); if (rcb) return rcb; }
    sum_double_kernel<<<blas_grid, BLAS_BLOCK, 0, stream>>>(cb ? tr_log_oo : tr_log, (size_t)g.Vh, make_red(0, blas_grid), scal + S_TMP0);
    int rc = launched("sum_double"); if (rc) return rc;
    rc = fetch_scalars(); if (rc) return rc;
    *out = h_scal[S_TMP0];
    return B200_OK;
  }

  // ------------------------------------------------------------------ fields
  int field_alloc(b200_field** f, int nrhs = 1) override {
    B200_CUDA(cudaSetDevice(cfg.device));
    if (nrhs < 1 || nrhs > MAX_SHIFT) { set_error("a batched field holds 1..%d vectors", MAX_SHIFT); return B200_ERR_ARG; }
    b200_field* p = new b200_field;
    p->bytes = sizeof(C) * nelem() * nrhs; p->prec = sizeof(R); p->d = nullptr; p->nrhs = nrhs;
    cudaError_t e = cudaMalloc(&p->d, p->bytes);
    if (e != cudaSuccess) { delete p; set_error("cudaMalloc(%zu) failed: %s", sizeof(C) * nelem() * nrhs, cudaGetErrorString(e)); return B200_ERR_CUDA; }
    cudaMemsetAsync(p->d, 0, p->bytes, stream);
    *f = p;
    return B200_OK;
  }
  void field_free(b200_field* f) override { if (f) { cudaSetDevice(cfg.device); cudaStreamSynchronize(stream); cudaFree(f->d); delete f; } }
  int field_zero(b200_field* f) override { B200_CUDA(cudaSetDevice(cfg.device)); B200_CUDA(cudaMemsetAsync(f->d, 0, f->bytes, stream)); return B200_OK; }
  int field_upload(b200_field* f, const void* host, int host_prec, int irhs = 0) override {
    B200_CUDA(cudaSetDevice(cfg.device));
    if (!f || !host) { set_error("null pointer"); return B200_ERR_ARG; }
    if (irhs < 0 || irhs >= f->nrhs) { set_error("right-hand side index %d out of range (field holds %d)", irhs, f->nrhs); return B200_ERR_ARG; }
    C* d = (C*)f->d + (size_t)irhs * nelem();
    if (host_prec == B200_DOUBLE) return upload_aos<double, 24, 12>((const double*)host, g.Vh, d, (size_t)g.Vh, MapIdentity(), 1.0);
    if (host_prec == B200_SINGLE) return upload_aos<float, 24, 12>((const float*)host, g.Vh, d, (size_t)g.Vh, MapIdentity(), 1.0);
    set_error("host_prec must be 4 or 8"); return B200_ERR_ARG;
  }
  int field_download(const b200_field* f, void* host, int host_prec, int irhs = 0) override {
    B200_CUDA(cudaSetDevice(cfg.device));
    if (!f || !host) { set_error("null pointer"); return B200_ERR_ARG; }
    if (irhs < 0 || irhs >= f->nrhs) { set_error("right-hand side index %d out of range (field holds %d)", irhs, f->nrhs); return B200_ERR_ARG; }
    const C* d = (const C*)f->d + (size_t)irhs * nelem();
    if (host_prec == B200_DOUBLE) return download_aos<double, 24, 12>((double*)host, g.Vh, d, (size_t)g.Vh, MapIdentity());
    if (host_prec == B200_SINGLE) return download_aos<float, 24, 12>((float*)host, g.Vh, d, (size_t)g.Vh, MapIdentity());
    set_error("host_prec must be 4 or 8"); return B200_ERR_ARG;
  }

  // workspace fields W(0..n-1), each able to hold `nrhs` right-hand sides
  int need_ws(int n, int nrhs = 1) {
    for (int i = 0; i < n; ++i) {
      if (ws[i] && ws[i]->nrhs < nrhs) { field_free(ws[i]); ws[i] = nullptr; }
      if (!ws[i]) { int rc = field_alloc(&ws[i], nrhs); if (rc) return rc; }
    }
    return B200_OK;
  }
  long l2_budget = 40l << 20;   // bytes of L2 the batched traversal may assume for its three live time slices (B200_MRHS_L2_KB overrides)
  int nb = 1;   // right-hand sides of the operation in flight (set by the public entry points; 1 = ordinary path)
  // Select the batch size and make sure the reduction scratch holds <= 4 partial sums per right-hand side and block,
  // for the BLAS grid as well as for the Dslash grids (32 sites per block when batched).
  int set_batch(int n) {
    if (n < 1 || n > MAX_RHS) { set_error("the batched kernels take 1..%d right-hand sides (field holds %d)", MAX_RHS, n); return B200_ERR_ARG; }
    nb = n;
    const size_t dblocks = n > 1 ? (size_t)(g.Vh + 31) / 32 + 4 : (size_t)(g.Vh + DSLASH_BLOCK - 1) / DSLASH_BLOCK + 8;
    return ensure_partial(4 * (size_t)n * std::max<size_t>((size_t)blas_grid, dblocks));
  }
  C* W(int i) { return (C*)ws[i]->d; }

  // ------------------------------------------------------------------ kernel launchers
  static constexpr int NRB = sizeof(R) == 4 ? B200_MRHS_NRB_F : B200_MRHS_NRB;   // right-hand sides per CTA (tuned per precision)
  template <int EPI>
  int launch_dslash(DslashArgs<R>& a) {
    a.gauge = gauge; a.scal = scal; a.status = status; a.g = g; a.pol = l2pol;
    a.nrhs = nb; a.fstride = nelem(); a.gstride = (size_t)6 * g.S3h; a.gstride_z = (size_t)6 * g.SZh;
    const int bs = nb > 1 ? 32 : DSLASH_BLOCK;          // target sites per CTA
    int rc;
    for (auto& b : a.box) b = SiteBox{0, 0, 0, 0};
    if (split()) {
      const SiteBox inner{g.tsplit ? 1 : 0, g.tsplit ? g.Lt - 2 : g.Lt, g.zsplit ? 1 : 0, g.zsplit ? g.Lz - 2 : g.Lz};
      const int n_int = box_count(g, inner);
      const int n_face = g.Vh - n_int;
      const int nb_int = (n_int + bs - 1) / bs;
      const int nb_face = (n_face + bs - 1) / bs;
      const int total = nb_int + nb_face;
      SiteBox faces[4]; int nf = 0;
      if (g.tsplit) { faces[nf++] = SiteBox{0, 1, 0, g.Lz}; faces[nf++] = SiteBox{g.Lt - 1, 1, 0, g.Lz}; }
      if (g.zsplit && inner.nt > 0) {
        faces[nf++] = SiteBox{inner.t0, inner.nt, 0, 1};
        faces[nf++] = SiteBox{inner.t0, inner.nt, g.Lz - 1, 1};
      }
      if (nb == 1) {
        // ONE launch: pack CTAs, interior CTAs, boundary CTAs that poll the arrival flags themselves (halo.cuh)
        HaloFuse<R> h;
        halo.prepare(h.pack, a.in, gauge, recon, ls, a.isign, a.parity, 1, a.fstride);
        for (int f = 0; f < 4; ++f) h.wait[f] = halo.local_flag(f);
        h.seq = halo.seq; h.spin = halo.spin_cycles;
        // one pack CTA per SM at most (they stride over the face sites): the other CTA slot of every SM starts interior
        // work at once, so the pack overlaps the interior instead of preceding it
        h.n_pack = std::min((halo.pack_threads() + DSLASH_BLOCK - 1) / DSLASH_BLOCK, num_sms); h.n_int = nb_int; h.n_int_sites = n_int;
        a.ghost_fwd = halo.ghost(0); a.ghost_bwd = halo.ghost(1); a.ghost_zfwd = halo.ghost(2); a.ghost_zbwd = halo.ghost(3);
        a.box[0] = inner; for (int k = 0; k < nf; ++k) a.box[1 + k] = faces[k];
        a.nbox = 1 + nf; a.nsites = g.Vh; set_chunks(a, 0); a.red = make_red(0, total);
        a.red.split = split_reduce<EPI>(total);
        rc = launch_halo<EPI>(a, h, h.n_pack + total); if (rc) return rc;
        return launch_finish<EPI>(a);
      }
      // batched right-hand sides: pack + send the faces over NVLink, run the interior while they fly, then the boundary
      rc = halo.start(a.in, gauge, recon, ls, a.isign, a.parity, a.check_stop ? status : nullptr, a.run_if, nb, a.fstride, launches); if (rc) return rc;
      a.ghost_fwd = halo.ghost(0); a.ghost_bwd = halo.ghost(1); a.ghost_zfwd = halo.ghost(2); a.ghost_zbwd = halo.ghost(3);
      const int split_b = split_reduce<EPI>(total);
      if (n_int > 0) {
        a.box[0] = inner; a.nbox = 1; a.nsites = n_int; a.red = make_red(0, total); a.red.split = split_b;
        set_chunks(a, inner.nz);
        rc = launch_one<EPI>(a, nb_int); if (rc) return rc;
      }
      rc = halo.wait(a.check_stop ? status : nullptr, a.run_if, nb, launches); if (rc) return rc;
      for (int k = 0; k < nf; ++k) a.box[k] = faces[k];
      a.nbox = nf; a.nsites = n_face; set_chunks(a, 0); a.red = make_red(nb_int, total); a.red.split = split_b;
      rc = launch_one<EPI>(a, nb_face); if (rc) return rc;
      return launch_finish<EPI>(a);
    }
    a.ghost_fwd = nullptr; a.ghost_bwd = nullptr; a.ghost_zfwd = nullptr; a.ghost_zbwd = nullptr;
    a.box[0] = SiteBox{0, g.Lt, 0, g.Lz}; a.nbox = 1; a.nsites = g.Vh;
    set_chunks(a, g.Lz);
    const int blocks = (g.Vh + bs - 1) / bs;
    a.red = make_red(0, blocks);
    a.red.split = split_reduce<EPI>(blocks);
    { int rc1 = launch_one<EPI>(a, blocks); if (rc1) return rc1; }
    return launch_finish<EPI>(a);
  }
  // Big single-RHS grids: the CTAs only store their partial sums, a one-CTA kernel behind the step finishes (reduce.cuh)
  int split_min_blocks = RED_FLAT_MAX;   // B200_SPLIT_MIN_BLOCKS overrides (tests: exercise the split path on small lattices)
  template <int EPI>
  int split_reduce(int blocks) const {
    return (EPI == EPI_M_NORM || EPI == EPI_M_CG || EPI == EPI_M_CGREL || EPI == EPI_M_DOTR0 || EPI == EPI_M_DOTX) && blocks > split_min_blocks;
  }
  template <int EPI>
  int launch_finish(const DslashArgs<R>& a) {
    if (!a.red.split) return B200_OK;
    if (nb > 1) {
      constexpr int E = (EPI == EPI_M_CGREL ? EPI_M_CG : EPI);     // (the reliable-update epilogue has no batched variant)
      dslash_mrhs_finish_kernel<R, E><<<nb, FINISH_BLOCK, 0, stream>>>(a);
      return launched("dslash_mrhs_finish_kernel");
    }
    dslash_finish_kernel<R, EPI><<<1, FINISH_BLOCK, 0, stream>>>(a);
    return launched("dslash_finish_kernel");
  }
  // the fused split-lattice launch (single right-hand side): dslash_halo_kernel, halo.cuh
  template <int EPI>
  int launch_halo(const DslashArgs<R>& a, const HaloFuse<R>& h, int blocks) {
    if (EPI >= EPI_M && a.mmode == MODE_SYM_PLUS) launch_halo_mode<EPI, (EPI >= EPI_M ? MODE_SYM_PLUS : MODE_ASYM)>(a, h, blocks);
    else if (EPI >= EPI_M && a.mmode == MODE_SYM_MINUS) launch_halo_mode<EPI, (EPI >= EPI_M ? MODE_SYM_MINUS : MODE_ASYM)>(a, h, blocks);
    else launch_halo_mode<EPI, MODE_ASYM>(a, h, blocks);
    return launched("dslash_halo_kernel");
  }
  template <int EPI, int MODE>
  void launch_halo_mode(const DslashArgs<R>& a, const HaloFuse<R>& h, int blocks) {
    const MrhsDiv dv = make_mrhs_div(a);
    if (recon == 12) dslash_halo_kernel<R, EPI, true, DSLASH_BLOCK, MODE><<<blocks, DSLASH_BLOCK, 0, stream>>>(a, ls, h, dv);
    else dslash_halo_kernel<R, EPI, false, DSLASH_BLOCK, MODE><<<blocks, DSLASH_BLOCK, 0, stream>>>(a, ls, h, dv);
  }
  template <int EPI>
  int launch_one(const DslashArgs<R>& a, int blocks) {
    if (blocks <= 0) return B200_OK;
    if (nb > 1) return launch_mrhs<EPI>(a, blocks);
    if (EPI >= EPI_M && a.mmode == MODE_SYM_PLUS) launch_mode<EPI, (EPI >= EPI_M ? MODE_SYM_PLUS : MODE_ASYM)>(a, blocks);
    else if (EPI >= EPI_M && a.mmode == MODE_SYM_MINUS) launch_mode<EPI, (EPI >= EPI_M ? MODE_SYM_MINUS : MODE_ASYM)>(a, blocks);
    else launch_mode<EPI, MODE_ASYM>(a, blocks);
    return launched("dslash_kernel");
  }
  template <int EPI, int MODE>
  void launch_mode(const DslashArgs<R>& a, int blocks) {
    const MrhsDiv dv = make_mrhs_div(a);
    if (recon == 12) dslash_kernel<R, EPI, true, DSLASH_BLOCK, MODE><<<blocks, DSLASH_BLOCK, 0, stream>>>(a, ls, dv);
    else dslash_kernel<R, EPI, false, DSLASH_BLOCK, MODE><<<blocks, DSLASH_BLOCK, 0, stream>>>(a, ls, dv);
  }
  // Traversal order of a box nz planes thick (launch_site / mrhs_site): chunks of dz z-planes x dy y-rows, all time slices
  // of a chunk before the next chunk, sized so that three time slices of a chunk (all right-hand sides) fit in
  // `l2_budget` (B200: 126 MB of L2, of which ~32 MB hold neighbour spinors in practice, profiles/r02_mrhs_zchunk_sweep.json).
  // Among the divisor pairs that fit, the one with the smallest surface wins: a chunk face is where a neighbour spinor is
  // fetched from DRAM a second time (2/dz + 2/dy extra fetches per site; a direction the chunk spans has no face).
  // nz = 0 or a box that fits whole: natural order (zc_sites = 0).
  void set_chunks(DslashArgs<R>& a, int nz) const {
    a.zc_sites = 0; a.zc_dz = nz; a.zc_dy = g.Ly; a.zc_ncy = 1;
    const long budget = l2_budget;
    const long rowb = (long)g.Lxh * 12 * (long)sizeof(C) * nb;              // one y-row of one time slice, all right-hand sides
    if (nz <= 0 || budget <= 0 || 3 * rowb * g.Ly * nz <= budget) return;
    int bz = 1, by = 1; double best = 1e30; long bvol = 0;
    for (int dz = 1; dz <= nz; ++dz) {
      if (nz % dz) continue;
      for (int dy = 1; dy <= g.Ly; ++dy) {
        if (g.Ly % dy || 3 * rowb * dz * dy > budget) continue;
        const double f = (dz < nz ? 2.0 / dz : 0.0) + (dy < g.Ly ? 2.0 / dy : 0.0);
        const long vol = (long)dz * dy;
        if (f < best - 1e-12 || (f < best + 1e-12 && vol > bvol)) { best = f; bz = dz; by = dy; bvol = vol; }
      }
    }
    if (!y_chunks) {        // B200_MRHS_YCHUNK=0: z-planes only (the round-1 order)
      by = g.Ly; bz = 1;
      for (int dz = 1; dz <= nz; ++dz) if (nz % dz == 0 && 3 * rowb * g.Ly * dz <= budget) bz = dz;
    }
    a.zc_dz = bz; a.zc_dy = by; a.zc_ncy = g.Ly / by;
    a.zc_sites = bz * by * g.Lxh;
  }
  int y_chunks = 1;
  // batched launch: CTA = 32 sites x NRB right-hand sides; the groups of one site block are adjacent in the grid
  template <int EPI>
  int launch_mrhs(const DslashArgs<R>& a, int site_blocks) {
    if (EPI == EPI_M_CGREL) { set_error("the reliable-update epilogue has no multi-RHS variant"); return B200_ERR_ARG; }
    constexpr int E = (EPI == EPI_M_CGREL ? EPI_M_CG : EPI);
    // the symmetric operator's epilogues (MODE_SYM_*) exist for the EPI_M* family only
    if (E >= EPI_M && a.mmode == MODE_SYM_PLUS) return launch_mrhs_mode<E, (E >= EPI_M ? MODE_SYM_PLUS : MODE_ASYM)>(a, site_blocks);
    if (E >= EPI_M && a.mmode == MODE_SYM_MINUS) return launch_mrhs_mode<E, (E >= EPI_M ? MODE_SYM_MINUS : MODE_ASYM)>(a, site_blocks);
    return launch_mrhs_mode<E, MODE_ASYM>(a, site_blocks);
  }
  template <int E, int MODE>
  int launch_mrhs_mode(const DslashArgs<R>& a, int site_blocks) {
    const int ngroups = (nb + NRB - 1) / NRB;
    const dim3 block(32, NRB);
    const MrhsDiv dv = make_mrhs_div(a);
    if (recon == 12) {
      auto k = dslash_mrhs_kernel<R, E, true, NRB, MODE>;
      B200_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MrhsSmem<R, E, true>::total(NRB)));
      k<<<site_blocks * ngroups, block, MrhsSmem<R, E, true>::total(NRB), stream>>>(a, ls, ngroups, dv);
    } else {
      auto k = dslash_mrhs_kernel<R, E, false, NRB, MODE>;
      B200_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MrhsSmem<R, E, false>::total(NRB)));
      k<<<site_blocks * ngroups, block, MrhsSmem<R, E, false>::total(NRB), stream>>>(a, ls, ngroups, dv);
    }
    return launched("dslash_mrhs_kernel");
  }

  int ready() {
    if (!gauge) { set_error("gauge field not loaded"); return B200_ERR_STATE; }
    if (!clov || !invclov) { set_error("clover term not loaded"); return B200_ERR_STATE; }
    if (sym && !invclov_oo) { set_error("symmetric preconditioning: A_oo^-1 not built"); return B200_ERR_STATE; }
    return B200_OK;
  }
  // workspace fields an entry point needs: the symmetric M^dag keeps A_oo^-1 x in W(8)
  int nws(int n) const { return sym ? 9 : n; }

  // out = M in (isign=+1) / M^dag in (-1) with one of the EPI_M* epilogues; te is the even temporary
  // Symmetric preconditioning (seoprec_clover_linop_w.cc:147-193):
  //   PLUS : t = A_ee^-1 D x ; out = x - 1/4 A_oo^-1 D t                (MODE_SYM_PLUS epilogue)
  //   MINUS: w = A_oo^-1 x ; t = A_ee^-1 D^dag w ; out = x - 1/4 D^dag t  (clover pass + MODE_SYM_MINUS epilogue)
  // (Measured and rejected: letting the PLUS epilogue also emit A_oo^-1 * out for the M^dag that follows in a
  //  normal-equation solver.  The second clover application can only load its matrix elements after m is known, and
  //  with 4 warps per SM that late L2 round trip costs more (+0.9 ms) than the stand-alone pass it saves (0.8 ms).)
  int apply_M(C* out, const C* in, int isign, int epi, C* r, const C* r0, int iter, int check, int run_if = 0) {
    const C* src = in;
    if (sym && isign < 0) {
      clover_kernel<R, 128><<<dim3((g.Vh + 127) / 128, nb), 128, 0, stream>>>(in, W(8), invclov_oo, g.Vh, nelem(), (check || run_if) ? status : nullptr, run_if);
      int rc0 = launched("clover_kernel"); if (rc0) return rc0;
      mark();
      src = W(8);
    }
    DslashArgs<R> a{};
    a.in = src; a.out = W(0); a.clov = invclov; a.parity = 0; a.isign = isign; a.iter = iter; a.check_stop = check; a.run_if = run_if;
    int rc = launch_dslash<EPI_AINV>(a); if (rc) return rc;
    mark();
    DslashArgs<R> b{};
    b.in = W(0); b.out = out; b.clov = clov + (size_t)36 * g.Vh; b.x = in; b.r = r; b.r0 = r0;
    if (sym) { b.clov = invclov_oo; b.mmode = isign > 0 ? MODE_SYM_PLUS : MODE_SYM_MINUS; }
    b.parity = 1; b.isign = isign; b.iter = iter; b.check_stop = check; b.run_if = run_if;
    b.twist = isign * twisted_m;
    switch (epi) {
      case EPI_M_CGREL: rc = launch_dslash<EPI_M_CGREL>(b); break;
      case EPI_M: rc = launch_dslash<EPI_M>(b); break;
      case EPI_M_NORM: rc = launch_dslash<EPI_M_NORM>(b); break;
      case EPI_M_CG: rc = launch_dslash<EPI_M_CG>(b); break;
      case EPI_M_DOTR0: rc = launch_dslash<EPI_M_DOTR0>(b); break;
      case EPI_M_DOTX: rc = launch_dslash<EPI_M_DOTX>(b); break;
      default: set_error("bad epilogue"); return B200_ERR_ARG;
    }
    mark();
    return rc;
  }

  int dslash(b200_field* out, const b200_field* in, int isign, int out_cb) override {
    B200_CUDA(cudaSetDevice(cfg.device));
    if (!gauge) { set_error("gauge field not loaded"); return B200_ERR_STATE; }
    if ((isign != 1 && isign != -1) || (out_cb != 0 && out_cb != 1) || !out || !in || out == in || out->nrhs != in->nrhs) { set_error("b200_dslash: bad argument"); return B200_ERR_ARG; }
    { int rcb = set_batch(in->nrhs); if (rcb) return rcb; }
    DslashArgs<R> a{};
    a.in = (const C*)in->d; a.out = (C*)out->d; a.parity = out_cb; a.isign = isign;
    return launch_dslash<EPI_DSLASH>(a);
  }

  int clover_apply(b200_field* out, const b200_field* in, int cb, int inverse) override {
    B200_CUDA(cudaSetDevice(cfg.device));
    if (!clov) { set_error("clover term not loaded"); return B200_ERR_STATE; }
    if ((cb != 0 && cb != 1) || !out || !in || out == in || out->nrhs != in->nrhs) { set_error("b200_clover_apply: bad argument"); return B200_ERR_ARG; }
    if (inverse && cb != 0 && !(sym && invclov_oo)) { set_error("without symmetric preconditioning only the cb-0 inverse exists (invclov.choles(0), eoprec_clover_linop_w.cc:30)"); return B200_ERR_ARG; }
    if (in->nrhs > MAX_RHS) { set_error("b200_clover_apply: at most %d right-hand sides", MAX_RHS); return B200_ERR_ARG; }
    const C* cl = inverse ? (cb ? invclov_oo : invclov) : clov + (size_t)cb * 36 * g.Vh;
    clover_kernel<R, 128><<<dim3((g.Vh + 127) / 128, in->nrhs), 128, 0, stream>>>((const C*)in->d, (C*)out->d, cl, g.Vh, nelem());
    return launched("clover_kernel");
  }

  int matpc(b200_field* out, const b200_field* in, int isign) override {
    B200_CUDA(cudaSetDevice(cfg.device));
    int rc = ready(); if (rc) return rc;
    if ((isign != 1 && isign != -1) || !out || !in || out == in || out->nrhs != in->nrhs) { set_error("b200_clover_matpc: bad argument"); return B200_ERR_ARG; }
    { int rcb = set_batch(in->nrhs); if (rcb) return rcb; }
    rc = need_ws(nws(1), nb); if (rc) return rc;
    return apply_M((C*)out->d, (const C*)in->d, isign, EPI_M, nullptr, nullptr, 0, 0);
  }

  int time_matpc(b200_field* out, const b200_field* in, int isign, int reps, double ms[2]) override {
    B200_CUDA(cudaSetDevice(cfg.device));
    int rc = ready(); if (rc) return rc;
    if ((isign != 1 && isign != -1) || !out || !in || out == in || reps < 1 || out->nrhs != in->nrhs) { set_error("b200_dev_time_matpc: bad argument"); return B200_ERR_ARG; }
    if (split()) { set_error("b200_dev_time_matpc: single-GPU measurement only"); return B200_ERR_ARG; }
    { int rcb = set_batch(in->nrhs); if (rcb) return rcb; }
    rc = need_ws(nws(1), nb); if (rc) return rc;
    cudaEvent_t e[3];
    for (auto& x : e) B200_CUDA(cudaEventCreate(&x));
    double acc[2] = {0, 0};
    for (int i = 0; i < reps; ++i) {
      DslashArgs<R> a{};
      a.in = (const C*)in->d; a.out = W(0); a.clov = invclov; a.parity = 0; a.isign = isign;
      DslashArgs<R> b{};
      b.in = W(0); b.out = (C*)out->d; b.clov = clov + (size_t)36 * g.Vh; b.x = (const C*)in->d; b.parity = 1; b.isign = isign;
      if (sym) { b.clov = invclov_oo; b.mmode = isign > 0 ? MODE_SYM_PLUS : MODE_SYM_MINUS; }
      b.twist = isign * twisted_m;
      B200_CUDA(cudaEventRecord(e[0], stream));
      if (sym && isign < 0) {   // the A_oo^-1 pass of the symmetric M^dag is booked with the first kernel
        clover_kernel<R, 128><<<dim3((g.Vh + 127) / 128, nb), 128, 0, stream>>>((const C*)in->d, W(8), invclov_oo, g.Vh, nelem());
        rc = launched("clover_kernel"); if (rc) return rc;
        a.in = W(8);
      }
      rc = launch_dslash<EPI_AINV>(a); if (rc) return rc;
      B200_CUDA(cudaEventRecord(e[1], stream));
      rc = launch_dslash<EPI_M>(b); if (rc) return rc;
      B200_CUDA(cudaEventRecord(e[2], stream));
      B200_CUDA(cudaEventSynchronize(e[2]));
      float t0 = 0, t1 = 0;
      B200_CUDA(cudaEventElapsedTime(&t0, e[0], e[1]));
      B200_CUDA(cudaEventElapsedTime(&t1, e[1], e[2]));
      acc[0] += t0; acc[1] += t1;
    }
    for (auto& x : e) cudaEventDestroy(x);
    ms[0] = acc[0] / reps; ms[1] = acc[1] / reps;
    return B200_OK;
  }

  int fetch_scalars() {
    B200_CUDA(cudaMemcpyAsync(h_scal, scal, sizeof(double) * S_COUNT * MAX_RHS, cudaMemcpyDeviceToHost, stream));
    B200_CUDA(cudaStreamSynchronize(stream));
    return comm_check();
  }
  double hs(int rhs, int slot) const { return h_scal[rhs * S_COUNT + slot]; }
  int set_scalars(const ScalarSet& s) {
    set_scalars_kernel<<<1, 32, 0, stream>>>(scal, status, s);
    return launched("set_scalars");
  }
  dim3 bgrid() const { return dim3(blas_grid, nb); }

  int norm2_dev(const C* x, int slot) {
    norm2_kernel<R><<<bgrid(), BLAS_BLOCK, 0, stream>>>(x, nelem(), make_red(0, blas_grid), scal + slot, nelem());
    return launched("norm2");
  }
  int xmy_norm_dev(C* out, C* out2, const C* x, const C* y, int slot) {
    xmy_norm_kernel<R><<<bgrid(), BLAS_BLOCK, 0, stream>>>(out, out2, x, y, nelem(), make_red(0, blas_grid), scal + slot, nelem());
    return launched("xmy_norm");
  }
  int axpby_dev(C* out, double a, const C* x, double b, const C* y) {
    axpby_kernel<R><<<bgrid(), BLAS_BLOCK, 0, stream>>>(out, a, x, b, y, nelem(), nelem());
    return launched("axpby");
  }

  int norm2(const b200_field* x, double* r) override {
    B200_CUDA(cudaSetDevice(cfg.device));
    { int rcb = set_batch(x->nrhs); if (rcb) return rcb; }
    int rc = norm2_dev((const C*)x->d, S_TMP0); if (rc) return rc;
    rc = fetch_scalars(); if (rc) return rc;
    for (int i = 0; i < nb; ++i) r[i] = hs(i, S_TMP0);
    return B200_OK;
  }
  int inner(const b200_field* x, const b200_field* y, double r[2]) override {
    B200_CUDA(cudaSetDevice(cfg.device));
    if (x->nrhs != y->nrhs) { set_error("b200_dev_inner: fields hold different numbers of right-hand sides"); return B200_ERR_ARG; }
    { int rcb = set_batch(x->nrhs); if (rcb) return rcb; }
    inner_kernel<R><<<bgrid(), BLAS_BLOCK, 0, stream>>>((const C*)x->d, (const C*)y->d, nelem(), make_red(0, blas_grid), scal + S_TMP0, nelem());
    int rc = launched("inner"); if (rc) return rc;
    rc = fetch_scalars(); if (rc) return rc;
    for (int i = 0; i < nb; ++i) { r[2 * i] = hs(i, S_TMP0); r[2 * i + 1] = hs(i, S_TMP1); }
    return B200_OK;
  }

  // ------------------------------------------------------------------ solver loops
  int ctl_rel = 0;   // set by the mixed-precision BiCGStab driver: bicg_update also takes the reliable-update decisions
  BlasCtl ctl(int iter, int check) {
    BlasCtl c; c.scal = scal; c.status = status; c.red = make_red(0, blas_grid); c.iter = iter; c.check_stop = check; c.fstride = nelem();
    c.rel = ctl_rel;
    return c;
  }

  // one CG iteration, invcg2.cc:158-220.  workspace: W(1)=mp, W(2)=p, W(3)=r
  int cg_iteration(C* psi, int k, int check) {
    int rc = apply_M(W(1), W(2), +1, EPI_M_NORM, nullptr, nullptr, k, check); if (rc) return rc;      // mp = M p, d, a
    rc = apply_M(nullptr, W(1), -1, EPI_M_CG, W(3), nullptr, k, check); if (rc) return rc;            // r -= a M^dag mp, cp, b
    cg_update_kernel<R><<<bgrid(), BLAS_BLOCK, 0, stream>>>(psi, W(2), W(3), nelem(), ctl(k, check));
    rc = launched("cg_update");
    mark();
    return rc;
  }
  // one BiCGStab iteration, invbicgstab.cc:74-170.  W(1)=r, W(2)=r0, W(3)=p, W(4)=v, W(5)=t
  int bicg_iteration(C* psi, int k, int check, int isign = +1) {
    bicg_p_kernel<R><<<bgrid(), BLAS_BLOCK, 0, stream>>>(W(3), W(1), W(4), nelem(), ctl(k, check));
    int rc = launched("bicg_p"); if (rc) return rc;
    mark();
    rc = apply_M(W(4), W(3), isign, EPI_M_DOTR0, nullptr, W(2), k, check); if (rc) return rc;         // v = M p, alpha
    bicg_s_kernel<R><<<bgrid(), BLAS_BLOCK, 0, stream>>>(W(1), W(4), nelem(), ctl(k, check));
    rc = launched("bicg_s"); if (rc) return rc;
    mark();
    rc = apply_M(W(5), W(1), isign, EPI_M_DOTX, nullptr, nullptr, k, check); if (rc) return rc;       // t = M r, omega
    bicg_update_kernel<R><<<bgrid(), BLAS_BLOCK, 0, stream>>>(psi, W(1), W(3), W(5), W(2), nelem(), ctl(k, check));
    rc = launched("bicg_update");
    mark();
    return rc;
  }

  // one multi-shift CG iteration, minvcg2.cc:243-342 (the p updates of :246-262 were done by the previous ms_update).
  // W(1)=Mp, W(2)=p0, W(3)=r; ms_p = p[s]; ms_psi = psi[s]
  static constexpr int SOLVER_MULTISHIFT = 100;
  int ms_iteration(int k, int check) {
    int rc = apply_M(W(1), W(2), +1, EPI_M_NORM, nullptr, nullptr, k, check); if (rc) return rc;      // d ; S_A = cp/d = -b
    rc = apply_M(nullptr, W(1), -1, EPI_M_CG, W(3), nullptr, k, check); if (rc) return rc;            // r += b M^dag M p0 ; c ; S_B = c/cp
    ms_scalars_kernel<<<1, 32, 0, stream>>>(ms_dev, scal, status, k, check);
    rc = launched("ms_scalars"); if (rc) return rc;
    ms_update_kernel<R><<<blas_grid, BLAS_BLOCK, 0, stream>>>(ms_psi, (C*)ms_p->d, W(2), W(3), nelem(), nelem(), ms_dev, status, k, check);
    return launched("ms_update");
  }

  // InvCG2_a set-up (invcg2.cc:100-150): returns chi_sq, cp in h_scal[S_TMP0], h_scal[S_TMP1]
  int cg_begin(C* psi, const C* chi) {
    int rc = norm2_dev(chi, S_TMP0); if (rc) return rc;
    rc = apply_M(W(1), psi, +1, EPI_M, nullptr, nullptr, 0, 0); if (rc) return rc;
    rc = apply_M(W(4), W(1), -1, EPI_M, nullptr, nullptr, 0, 0); if (rc) return rc;
    rc = xmy_norm_dev(W(3), W(2), chi, W(4), S_TMP1); if (rc) return rc;     // r = chi - M^dag M psi ; p = r ; cp
    return fetch_scalars();
  }
  // InvBiCGStab_a set-up (invbicgstab.cc:31-71)
  int bicg_begin(C* psi, const C* chi, int isign = +1) {
    int rc = norm2_dev(chi, S_TMP0); if (rc) return rc;
    rc = apply_M(W(2), psi, isign, EPI_M, nullptr, nullptr, 0, 0); if (rc) return rc;
    rc = xmy_norm_dev(W(1), W(2), chi, W(2), S_TMP1); if (rc) return rc;     // r = r0 = chi - M psi ; |r|^2 = <r0|r>
    B200_CUDA(cudaMemsetAsync(W(3), 0, sizeof(C) * nelem() * nb, stream));
    B200_CUDA(cudaMemsetAsync(W(4), 0, sizeof(C) * nelem() * nb, stream));
    return fetch_scalars();
  }

  // Enqueue iterations in batches of ITER_BATCH and poll the status blocks with a pipelined async copy; stop when
  // every right-hand side has either converged or broken down.  Outputs are arrays of nb entries.
  int poll_loop(C* psi, int solver, int max_iter, int* n_count, int* converged, int* breakdown, int isign = +1) {
    int k = 1, slot = 0, prev = -1;
    bool done = false;
    const int SB = ST_COUNT * MAX_RHS;
    while (k <= max_iter && !done) {
      const int n = std::min(ITER_BATCH, max_iter - k + 1);
      for (int i = 0; i < n; ++i) {
        int rc = (solver == B200_SOLVER_CG) ? cg_iteration(psi, k + i, 1)
               : (solver == SOLVER_MULTISHIFT) ? ms_iteration(k + i, 1) : bicg_iteration(psi, k + i, 1, isign);
        if (rc) return rc;
      }
      B200_CUDA(cudaMemcpyAsync(h_status + slot * SB, status, sizeof(int) * SB, cudaMemcpyDeviceToHost, stream));
      B200_CUDA(cudaEventRecord(ev_poll[slot], stream));
      if (prev >= 0) {
        B200_CUDA(cudaEventSynchronize(ev_poll[prev]));
        done = true;
        for (int r = 0; r < nb; ++r)
          if (h_status[prev * SB + r * ST_COUNT + ST_STOP] == 0 && h_status[prev * SB + r * ST_COUNT + ST_BREAKDOWN] == 0) done = false;
      }
      prev = slot; slot ^= 1; k += n;
    }
    B200_CUDA(cudaStreamSynchronize(stream));
    if (prev < 0) {   // max_iter == 0: nothing ran
      for (int r = 0; r < nb; ++r) { breakdown[r] = 0; converged[r] = converged[r] ? 1 : 0; n_count[r] = 0; }
      return B200_OK;
    }
    for (int r = 0; r < nb; ++r) {
      const int* st = h_status + prev * SB + r * ST_COUNT;   // the last copy enqueued reflects the final state
      breakdown[r] = st[ST_BREAKDOWN];
      converged[r] = st[ST_STOP] != 0;
      n_count[r] = st[ST_STOP] > 0 ? st[ST_STOP] : (st[ST_STOP] < 0 ? 0 : max_iter);   // -1 = converged before iterating
    }
    return B200_OK;
  }

  // InvCG2_a (invcg2.cc:70-232): solve M^dag M psi = rhs for nb right-hand sides in lockstep.  Uses W(0..4).
  int run_cg(C* psi, const C* rhs, double rsd, int max_iter, int* n_count, int* converged, double* rsd_sq_iter) {
    int rc = cg_begin(psi, rhs); if (rc) return rc;
    bool any = false;
    for (int r = 0; r < nb; ++r) {
      const double chi_sq = hs(r, S_TMP0), cp = hs(r, S_TMP1), rsd_sq = rsd * rsd * chi_sq;
      rsd_sq_iter[r] = cp; n_count[r] = 0; converged[r] = cp <= rsd_sq;          // invcg2.cc:136-146
      ScalarSet s{}; s.rhs = r; s.n = 2; s.slots[0] = S_RSDSQ; s.vals[0] = rsd_sq; s.slots[1] = S_C; s.vals[1] = cp;
      s.reset_status = converged[r] ? 2 : 1;                                      // 2: reset, then mark "stopped at iteration 0"
      rc = set_scalars(s); if (rc) return rc;
      any = any || !converged[r];
    }
    if (!any) return B200_OK;
    int breakdown[MAX_RHS];
    rc = poll_loop(psi, B200_SOLVER_CG, max_iter, n_count, converged, breakdown); if (rc) return rc;
    rc = fetch_scalars(); if (rc) return rc;
    for (int r = 0; r < nb; ++r) {
      if (n_count[r] > 0) rsd_sq_iter[r] = hs(r, S_CP);
      if (breakdown[r] >= 90) return comm_timeout(breakdown[r]);
    }
    return B200_OK;
  }
  // InvBiCGStab_a (invbicgstab.cc:10-202): solve M psi = rhs (isign=+1) or M^dag psi = rhs (-1).  Uses W(0..5).
  int run_bicg(C* psi, const C* rhs, int isign, double rsd, int max_iter, int* n_count, int* converged, double* rsd_sq_iter) {
    int rc = bicg_begin(psi, rhs, isign); if (rc) return rc;
    int breakdown[MAX_RHS];
    bool any = false;
    for (int r = 0; r < nb; ++r) {
      const double chi_sq = hs(r, S_TMP0), rr = hs(r, S_TMP1), rsd_sq = rsd * rsd * chi_sq;
      rsd_sq_iter[r] = rr; n_count[r] = 0; converged[r] = 0; breakdown[r] = 0;
      ScalarSet s{}; s.rhs = r; s.reset_status = 1; s.n = 11;
      const int sl[11] = {S_RSDSQ, S_RHO_RE, S_RHO_IM, S_RHOP_RE, S_RHOP_IM, S_ALPHA_RE, S_ALPHA_IM, S_OMEGA_RE, S_OMEGA_IM, S_BETA_RE, S_BETA_IM};
      const double vl[11] = {rsd_sq, rr, 0.0, 1.0, 0.0, 1.0, 0.0, 1.0, 0.0, rr, 0.0};   // beta_1 = (rho_1/1)(1/1)
      for (int i = 0; i < 11; ++i) { s.slots[i] = sl[i]; s.vals[i] = vl[i]; }
      if (rr == 0.0) { breakdown[r] = 1; s.reset_status = 3; }                    // rho = <r0|r> = 0 (invbicgstab.cc:80-83)
      rc = set_scalars(s); if (rc) return rc;
      any = any || rr != 0.0;
    }
    if (any) {
      rc = poll_loop(psi, B200_SOLVER_BICGSTAB, max_iter, n_count, converged, breakdown, isign); if (rc) return rc;
      rc = fetch_scalars(); if (rc) return rc;
    }
    int bad = 0;
    for (int r = 0; r < nb; ++r) {
      if (n_count[r] > 0) rsd_sq_iter[r] = hs(r, S_RNORM);
      if (breakdown[r] >= 90) return comm_timeout(breakdown[r]);
      if (breakdown[r]) { bad = breakdown[r]; set_error("BiCGStab breakdown (code %d) on right-hand side %d at iteration <= %d", bad, r, n_count[r]); }
    }
    return bad ? B200_ERR_BREAKDOWN : B200_OK;
  }
  int comm_timeout(int code) {
    set_error("multi-GPU peer wait timed out (code %d: 90 = halo flag, 91 = reduction mailbox)", code);
    return B200_ERR_COMM;
  }

  // True residual of the shells: |chi - M psi| (syssolver_linop_cg.h:80-87) or |chi - M^dag M psi| (syssolver_mdagm_cg.h:75-82)
  int true_residual(const C* psi, const C* chi, int mdagm, b200_solve_info* info) {
    int rc = apply_M(W(1), psi, +1, EPI_M, nullptr, nullptr, 0, 0); if (rc) return rc;
    const C* mpsi = W(1);
    if (mdagm) { rc = apply_M(W(2), W(1), -1, EPI_M, nullptr, nullptr, 0, 0); if (rc) return rc; mpsi = W(2); }
    rc = xmy_norm_dev(nullptr, nullptr, chi, mpsi, S_TMP2); if (rc) return rc;
    rc = norm2_dev(chi, S_TMP3); if (rc) return rc;
    rc = fetch_scalars(); if (rc) return rc;
    for (int r = 0; r < nb; ++r) {
      info[r].resid = sqrt(hs(r, S_TMP2));
      info[r].rel_resid = hs(r, S_TMP3) > 0 ? info[r].resid / sqrt(hs(r, S_TMP3)) : 0.0;
    }
    return B200_OK;
  }

  // psi, chi may hold several right-hand sides (b200_mfield_alloc): they are then solved in lockstep by the batched
  // kernels, each with its own scalars and stopping test; info is an array of psi->nrhs entries.
  int invert(b200_field* psi_f, const b200_field* chi_f, int solver, double rsd, int max_iter, int mdagm, b200_solve_info* info) override {
    B200_CUDA(cudaSetDevice(cfg.device));
    int rc = ready(); if (rc) return rc;
    if (!psi_f || !chi_f || !info || psi_f == chi_f || max_iter < 0 || !(rsd >= 0.0) || psi_f->nrhs != chi_f->nrhs) { set_error("b200_invert: bad argument"); return B200_ERR_ARG; }
    if (solver != B200_SOLVER_CG && solver != B200_SOLVER_BICGSTAB) { set_error("unknown solver %d", solver); return B200_ERR_ARG; }
    { int rcb = set_batch(psi_f->nrhs); if (rcb) return rcb; }
    rc = need_ws(nws(7), nb); if (rc) return rc;
    C* psi = (C*)psi_f->d; const C* chi = (const C*)chi_f->d;
    memset(info, 0, sizeof(*info) * nb);
    B200_CUDA(cudaEventRecord(ev_t0, stream));
    int n_count[MAX_RHS] = {0}, converged[MAX_RHS] = {0};
    double flops_iter, rsq[MAX_RHS] = {0.0};
    if (solver == B200_SOLVER_CG) {
      flops_iter = 2.0 * 3792.0 + 240.0;
      const C* rhs = chi;
      if (!mdagm) {   // chi_tmp = M^dag chi (syssolver_linop_cg.h:65-66) -> W(6)
        rc = apply_M(W(6), chi, -1, EPI_M, nullptr, nullptr, 0, 0); if (rc) return rc;
        rhs = W(6);
      }
      rc = run_cg(psi, rhs, rsd, max_iter, n_count, converged, rsq);
    } else if (!mdagm) {
      flops_iter = 2.0 * 3792.0 + 960.0;
      rc = run_bicg(psi, chi, +1, rsd, max_iter, n_count, converged, rsq);
    } else {
      // two-step solve, syssolver_mdagm_bicgstab.h:62-110: Y = M psi; M^dag Y = chi; M psi = Y
      flops_iter = 2.0 * 3792.0 + 960.0;
      int n1[MAX_RHS] = {0}, c1[MAX_RHS] = {0};
      rc = apply_M(W(6), psi, +1, EPI_M, nullptr, nullptr, 0, 0);
      if (!rc) rc = run_bicg(W(6), chi, -1, rsd, max_iter, n1, c1, rsq);
      if (!rc) rc = run_bicg(psi, W(6), +1, rsd, max_iter, n_count, converged, rsq);
      for (int r = 0; r < nb; ++r) { n_count[r] += n1[r]; converged[r] = converged[r] && c1[r]; }
    }
    for (int r = 0; r < nb; ++r) { info[r].n_count = n_count[r]; info[r].converged = converged[r]; info[r].rsd_sq_iter = rsq[r]; }
    if (rc && rc != B200_ERR_BREAKDOWN) return rc;
    B200_CUDA(cudaEventRecord(ev_t1, stream));
    int rc2 = true_residual(psi, chi, mdagm, info); if (rc2) return rc2;
    float ms = 0.f;
    B200_CUDA(cudaEventElapsedTime(&ms, ev_t0, ev_t1));
    const double gvol = (double)g.Vh * nranks();
    for (int r = 0; r < nb; ++r) {
      info[r].secs = ms * 1e-3; info[r].secs_total = info[r].secs;       // the batch shares one wall clock
      info[r].gflops = ms > 0 ? flops_iter * gvol * n_count[r] / (ms * 1e-3) * 1e-9 : 0.0;
    }
    return rc;
  }

  // MInvCG2_a (minvcg2.cc:74-373) behind MdagMMultiSysSolverCG::operator() (multi_syssolver_mdagm_cg.h:58-105):
  // (M^dag M + shifts[s]) psi[s] = chi.  psi_f is a batched field with at least n_shift vectors (zeroed here, as the
  // reference does); info[s] gets the common iteration count and the TRUE relative residual of shift s.
  int invert_multishift(b200_field* psi_f, const b200_field* chi_f, int n_shift, const double* shifts, const double* rsd,
                        int max_iter, b200_solve_info* info) override {
    B200_CUDA(cudaSetDevice(cfg.device));
    int rc = ready(); if (rc) return rc;
    if (!psi_f || !chi_f || !shifts || !rsd || !info || psi_f == chi_f || max_iter < 0) { set_error("b200_invert_multishift: bad argument"); return B200_ERR_ARG; }
    if (n_shift < 1 || n_shift > MAX_SHIFT) { set_error("b200_invert_multishift: 1..%d shifts (got %d)", MAX_SHIFT, n_shift); return B200_ERR_ARG; }
    if (psi_f->nrhs < n_shift || chi_f->nrhs != 1) { set_error("b200_invert_multishift: psi must hold n_shift vectors, chi one"); return B200_ERR_ARG; }
    for (int s = 0; s < n_shift; ++s) if (!(rsd[s] >= 0.0)) { set_error("b200_invert_multishift: negative residual target"); return B200_ERR_ARG; }
    { int rcb = set_batch(1); if (rcb) return rcb; }
    rc = need_ws(nws(5), 1); if (rc) return rc;
    if (ms_p && ms_p->nrhs < n_shift) { field_free(ms_p); ms_p = nullptr; }
    if (!ms_p) { rc = field_alloc(&ms_p, n_shift); if (rc) return rc; }
    if (!ms_dev) B200_CUDA(cudaMalloc(&ms_dev, sizeof(MsState)));
    C* psi = (C*)psi_f->d; const C* chi = (const C*)chi_f->d;
    ms_psi = psi;
    const size_t vbytes = sizeof(C) * nelem();
    memset(info, 0, sizeof(*info) * n_shift);
    B200_CUDA(cudaEventRecord(ev_t0, stream));
    B200_CUDA(cudaMemsetAsync(psi, 0, vbytes * n_shift, stream));                    // minvcg2.cc:118-122
    rc = norm2_dev(chi, S_TMP0); if (rc) return rc;
    rc = fetch_scalars(); if (rc) return rc;
    const double chi_norm_sq = hs(0, S_TMP0);
    int n_count = 0, converged = 1;
    if (!(sqrt(chi_norm_sq) < 1.0e-5)) {                                             // fuzz, minvcg2.cc:135-148
      MsState h;
      memset(&h, 0, sizeof(h));
      h.n_shift = n_shift; h.isz = 0;
      for (int s = 0; s < n_shift; ++s) {
        h.shift[s] = shifts[s]; h.rsd_sq[s] = chi_norm_sq * rsd[s] * rsd[s];
        if (shifts[s] < shifts[h.isz]) h.isz = s;
      }
      B200_CUDA(cudaMemcpyAsync(ms_dev, &h, sizeof(h), cudaMemcpyHostToDevice, stream));
      B200_CUDA(cudaStreamSynchronize(stream));                                      // h lives on this stack frame
      B200_CUDA(cudaMemcpyAsync(W(3), chi, vbytes, cudaMemcpyDeviceToDevice, stream));   // r = p0 = p[s] = chi (:166-179)
      B200_CUDA(cudaMemcpyAsync(W(2), chi, vbytes, cudaMemcpyDeviceToDevice, stream));
      for (int s = 0; s < n_shift; ++s)
        B200_CUDA(cudaMemcpyAsync((C*)ms_p->d + (size_t)s * nelem(), chi, vbytes, cudaMemcpyDeviceToDevice, stream));
      ScalarSet sc{}; sc.rhs = 0; sc.reset_status = 1; sc.n = 2;
      sc.slots[0] = S_RSDSQ; sc.vals[0] = -1.0;                                      // the fused |r|^2 finaliser must never stop the solve
      sc.slots[1] = S_C; sc.vals[1] = chi_norm_sq;                                   // cp = |chi|^2
      rc = set_scalars(sc); if (rc) return rc;
      // b = -cp/d ; r += b M^dag M p0 ; z, bs ; psi[s] = -bs chi ; c = |r|^2       (:186-232) -- iteration "0"
      rc = ms_iteration(0, 0); if (rc) return rc;
      rc = fetch_scalars(); if (rc) return rc;
      const double c0 = hs(0, S_CP);
      converged = c0 < h.rsd_sq[h.isz];                                              // :240
      if (!converged) {
        int nc[MAX_RHS] = {0}, cv[MAX_RHS] = {0}, bd[MAX_RHS] = {0};
        rc = poll_loop(psi, SOLVER_MULTISHIFT, max_iter, nc, cv, bd); if (rc) return rc;
        if (bd[0] >= 90) return comm_timeout(bd[0]);
        n_count = nc[0]; converged = cv[0];
      }
    }
    B200_CUDA(cudaEventRecord(ev_t1, stream));
    // per-shift true residuals, multi_syssolver_mdagm_cg.h:84-99
    MsState hfin;
    memset(&hfin, 0, sizeof(hfin));
    if (n_count > 0) { B200_CUDA(cudaMemcpyAsync(&hfin, ms_dev, sizeof(hfin), cudaMemcpyDeviceToHost, stream)); B200_CUDA(cudaStreamSynchronize(stream)); }
    for (int s = 0; s < n_shift; ++s) {
      const C* ps = psi + (size_t)s * nelem();
      rc = apply_M(W(1), ps, +1, EPI_M, nullptr, nullptr, 0, 0); if (rc) return rc;
      rc = apply_M(W(4), W(1), -1, EPI_M, nullptr, nullptr, 0, 0); if (rc) return rc;
      shifted_resid_kernel<R><<<blas_grid, BLAS_BLOCK, 0, stream>>>(chi, W(4), ps, shifts[s], nelem(), make_red(0, blas_grid), scal + S_TMP2);
      rc = launched("shifted_resid"); if (rc) return rc;
      rc = fetch_scalars(); if (rc) return rc;
      info[s].resid = sqrt(hs(0, S_TMP2));
      info[s].rel_resid = chi_norm_sq > 0 ? info[s].resid / sqrt(chi_norm_sq) : 0.0;
      info[s].n_count = n_count; info[s].converged = converged;
      info[s].rsd_sq_iter = n_count > 0 ? hfin.css[s] : 0.0;
    }
    float ms_t = 0.f;
    B200_CUDA(cudaEventElapsedTime(&ms_t, ev_t0, ev_t1));
    const double gvol = (double)g.Vh * nranks();
    // flops as MInvCG2_a books them: 2 M + 4*(4*Nc*Ns) for p0, r and the two norms + (6+2)*Nc*Ns per shift (minvcg2.cc:250-305)
    const double flops_iter = 2.0 * 3792.0 + 4.0 * 48.0 + n_shift * 96.0;
    for (int s = 0; s < n_shift; ++s) {
      info[s].secs = ms_t * 1e-3; info[s].secs_total = info[s].secs;
      info[s].gflops = ms_t > 0 ? flops_iter * gvol * n_count / (ms_t * 1e-3) * 1e-9 : 0.0;
    }
    return B200_OK;
  }

  int iterate_begin(b200_field* psi_f, const b200_field* chi_f, int solver) override {
    B200_CUDA(cudaSetDevice(cfg.device));
    int rc = ready(); if (rc) return rc;
    if (psi_f->nrhs != chi_f->nrhs) { set_error("b200_dev_iterate_begin: fields hold different numbers of right-hand sides"); return B200_ERR_ARG; }
    { int rcb = set_batch(psi_f->nrhs); if (rcb) return rcb; }
    rc = need_ws(nws(7), nb); if (rc) return rc;
    it_psi = (C*)psi_f->d; it_chi = (const C*)chi_f->d; it_k = 0; it_nb = nb;
    if (solver == B200_SOLVER_CG) { rc = cg_begin(it_psi, it_chi); if (rc) return rc; }
    else { rc = bicg_begin(it_psi, it_chi, +1); if (rc) return rc; }
    for (int r = 0; r < nb; ++r) {
      ScalarSet s{}; s.reset_status = 1; s.rhs = r;
      if (solver == B200_SOLVER_CG) {
        s.n = 2; s.slots[0] = S_RSDSQ; s.vals[0] = 0.0; s.slots[1] = S_C; s.vals[1] = hs(r, S_TMP1);
      } else {
        const double rr = hs(r, S_TMP1);
        s.n = 11;
        const int sl[11] = {S_RSDSQ, S_RHO_RE, S_RHO_IM, S_RHOP_RE, S_RHOP_IM, S_ALPHA_RE, S_ALPHA_IM, S_OMEGA_RE, S_OMEGA_IM, S_BETA_RE, S_BETA_IM};
        const double vl[11] = {0.0, rr, 0.0, 1.0, 0.0, 1.0, 0.0, 1.0, 0.0, rr, 0.0};
        for (int i = 0; i < 11; ++i) { s.slots[i] = sl[i]; s.vals[i] = vl[i]; }
      }
      rc = set_scalars(s); if (rc) return rc;
    }
    return B200_OK;
  }
  int iterate(int solver, int n_iter) override {
    B200_CUDA(cudaSetDevice(cfg.device));
    if (!it_psi) { set_error("b200_dev_iterate: call b200_dev_iterate_begin first"); return B200_ERR_STATE; }
    { int rcb = set_batch(it_nb); if (rcb) return rcb; }
    for (int i = 0; i < n_iter; ++i) {
      ++it_k;
      int rc = (solver == B200_SOLVER_CG) ? cg_iteration(it_psi, it_k, 0) : bicg_iteration(it_psi, it_k, 0);
      if (rc) return rc;
    }
    return B200_OK;
  }

  // Measurement aid behind b200_dev_time_solver_kernels: `reps` iterations of the loop set up by iterate_begin, with an
  // event behind every launch; ms[i] = average duration of the i-th launch of one iteration (CG: EPI_AINV, EPI_M_NORM,
  // EPI_AINV^dag, EPI_M_CG, cg_update; BiCGStab: bicg_p, EPI_AINV, EPI_M_DOTR0, bicg_s, EPI_AINV, EPI_M_DOTX, bicg_update).
  int time_solver_kernels(int solver, int reps, double* ms, int max_ms, int* n_ms) override {
    B200_CUDA(cudaSetDevice(cfg.device));
    if (!it_psi) { set_error("b200_dev_time_solver_kernels: call b200_dev_iterate_begin first"); return B200_ERR_STATE; }
    if (reps < 1 || !ms || !n_ms || max_ms < 1) { set_error("b200_dev_time_solver_kernels: bad argument"); return B200_ERR_ARG; }
    { int rcb = set_batch(it_nb); if (rcb) return rcb; }
    if (trace_ev.empty()) { trace_ev.resize(17); for (auto& e : trace_ev) B200_CUDA(cudaEventCreate(&e)); }
    std::vector<double> acc(16, 0.0);
    int n = 0;
    for (int i = 0; i < reps; ++i) {
      ++it_k;
      trace_n = 0; mark();
      int rc = (solver == B200_SOLVER_CG) ? cg_iteration(it_psi, it_k, 0) : bicg_iteration(it_psi, it_k, 0);
      n = trace_n - 1; trace_n = -1;
      if (rc) return rc;
      B200_CUDA(cudaEventSynchronize(trace_ev[n]));
      for (int k = 0; k < n; ++k) { float t = 0.f; B200_CUDA(cudaEventElapsedTime(&t, trace_ev[k], trace_ev[k + 1])); acc[k] += t; }
    }
    *n_ms = n;
    for (int k = 0; k < n && k < max_ms; ++k) ms[k] = acc[k] / reps;
    return comm_check();
  }

  // ------------------------------------------------------------------ full-lattice propagator (section 8f, rank 1)
  // PrecFermActQprop::operator() (eoprec_fermact_qprop.cc:41-80) inside the 12-colour-spin loop of
  // quarkProp4_a (quarkprop4_w.cc:70-117), all on the device.  evenOddLinOp = -1/2 D (eoprec_clover_linop_w.cc:98-133).
  int qprop(void* psi_h, const void* chi_h, int host_prec, int nrhs, int solver, double rsd, int max_iter,
            b200_solve_info* infos) override {
    B200_CUDA(cudaSetDevice(cfg.device));
    int rc = ready(); if (rc) return rc;
    if (!psi_h || !chi_h || !infos || nrhs < 1) { set_error("b200_qprop: bad argument"); return B200_ERR_ARG; }
    if (host_prec != B200_SINGLE && host_prec != B200_DOUBLE) { set_error("host_prec must be 4 or 8"); return B200_ERR_ARG; }
    // How many right-hand sides go through the batched kernels at once: all of them (up to MAX_RHS) if the 12 batched
    // vectors the solve needs (7 workspace + 5 here) fit in free HBM, else the largest batch that does.
    size_t free_b = 0, total_b = 0;
    B200_CUDA(cudaMemGetInfo(&free_b, &total_b));
    for (auto f : ws) if (f) free_b += f->bytes;
    const size_t per_rhs = (size_t)12 * sizeof(C) * nelem() + (size_t)3 * sizeof(double) * (g.Vh / 32 + 8);
    int cap = (int)std::min<size_t>((size_t)MAX_RHS, (size_t)(0.9 * (double)free_b) / per_rhs);
    if (const char* e = getenv("B200_QPROP_BATCH")) cap = std::min(cap, std::max(1, atoi(e)));
    if (cap < 1) { set_error("b200_qprop: not enough free device memory for even one right-hand side"); return B200_ERR_CUDA; }
    cap = std::min(cap, nrhs);
    const size_t cbbytes = (size_t)g.Vh * 24 * host_prec;
    b200_field *chi_e = nullptr, *chi_o = nullptr, *psi_o = nullptr, *t1 = nullptr, *t2 = nullptr;
    if ((rc = field_alloc(&chi_e, cap)) || (rc = field_alloc(&chi_o, cap)) || (rc = field_alloc(&psi_o, cap)) ||
        (rc = field_alloc(&t1, cap)) || (rc = field_alloc(&t2, cap))) {
      field_free(chi_e); field_free(chi_o); field_free(psi_o); field_free(t1); field_free(t2);
      return rc;
    }
    int worst = B200_OK;
    for (int i0 = 0; i0 < nrhs && !rc; i0 += cap) {
      const int n = std::min(cap, nrhs - i0);
      chi_e->nrhs = chi_o->nrhs = psi_o->nrhs = t1->nrhs = t2->nrhs = n;      // views of the first n right-hand sides
      for (int j = 0; j < n && !rc; ++j) {
        const char* ch = (const char*)chi_h + (size_t)(i0 + j) * 2 * cbbytes;
        const char* ph = (const char*)psi_h + (size_t)(i0 + j) * 2 * cbbytes;
        if ((rc = field_upload(chi_e, ch, host_prec, j))) break;
        if ((rc = field_upload(chi_o, ch + cbbytes, host_prec, j))) break;
        rc = field_upload(psi_o, ph + cbbytes, host_prec, j);
      }
      if (rc) break;
      if (sym) {
        // SymEvenOddPrecActQprop::operator(), seoprec_fermact_qprop.cc:45-89 with M_oe = A_oo^-1 (-1/2 Dslash),
        // M_eo = A_ee^-1 (-1/2 Dslash) (lib/seoprec_linop.h:170-204), for the whole batch:
        //   chi'_e = A_ee^-1 chi_e ; chi'_o = A_oo^-1 (chi_o + 1/2 Dslash chi'_e) ; S psi_o = chi'_o ;
        //   psi_e = chi'_e + 1/2 A_ee^-1 Dslash psi_o
        if ((rc = clover_apply(t1, chi_e, 0, 1))) break;
        if ((rc = dslash(t2, t1, +1, 1))) break;
        if ((rc = set_batch(n))) break;
        if ((rc = axpby_dev((C*)chi_o->d, 1.0, (const C*)chi_o->d, 0.5, (const C*)t2->d))) break;
        if ((rc = clover_apply(t2, chi_o, 1, 1))) break;
        rc = invert(psi_o, t2, solver, rsd, max_iter, 0, &infos[i0]);
        if (rc == B200_ERR_BREAKDOWN) { worst = rc; rc = B200_OK; }
        if (rc) break;
        if ((rc = dslash(chi_o, psi_o, +1, 0))) break;
        if ((rc = clover_apply(t2, chi_o, 0, 1))) break;
        if ((rc = set_batch(n))) break;
        if ((rc = axpby_dev((C*)chi_e->d, 1.0, (const C*)t1->d, 0.5, (const C*)t2->d))) break;
        for (int j = 0; j < n && !rc; ++j) {
          char* ph = (char*)psi_h + (size_t)(i0 + j) * 2 * cbbytes;
          if ((rc = field_download(chi_e, ph, host_prec, j))) break;
          rc = field_download(psi_o, ph + cbbytes, host_prec, j);
        }
        continue;
      }
      // chi' = chi_o - D_oe A_ee^-1 chi_e = chi_o + 1/2 Dslash(A_ee^-1 chi_e)
      if ((rc = clover_apply(t1, chi_e, 0, 1))) break;
      if ((rc = dslash(t2, t1, +1, 1))) break;
      if ((rc = set_batch(n))) break;
      if ((rc = axpby_dev((C*)t1->d, 1.0, (const C*)chi_o->d, 0.5, (const C*)t2->d))) break;
      rc = invert(psi_o, t1, solver, rsd, max_iter, 0, &infos[i0]);
      if (rc == B200_ERR_BREAKDOWN) { worst = rc; rc = B200_OK; }
      if (rc) break;
      // psi_e = A_ee^-1 (chi_e - D_eo psi_o) = A_ee^-1 (chi_e + 1/2 Dslash psi_o)
      if ((rc = dslash(t1, psi_o, +1, 0))) break;
      if ((rc = set_batch(n))) break;
      if ((rc = axpby_dev((C*)t2->d, 1.0, (const C*)chi_e->d, 0.5, (const C*)t1->d))) break;
      if ((rc = clover_apply(t1, t2, 0, 1))) break;
      for (int j = 0; j < n && !rc; ++j) {
        char* ph = (char*)psi_h + (size_t)(i0 + j) * 2 * cbbytes;
        if ((rc = field_download(t1, ph, host_prec, j))) break;
        rc = field_download(psi_o, ph + cbbytes, host_prec, j);
      }
    }
    chi_e->nrhs = chi_o->nrhs = psi_o->nrhs = t1->nrhs = t2->nrhs = cap;
    field_free(chi_e); field_free(chi_o); field_free(psi_o); field_free(t1); field_free(t2);
    return rc ? rc : worst;
  }
};

}  // namespace b200
