// engine.cuh -- host-side engine: device state, kernel launchers and the CG / BiCGStab drivers.
//
// Everything the plugin needs after construction happens on one CUDA stream; solver scalars never
// leave the device, the host only polls a status word every ITER_BATCH iterations (pipelined so the
// poll never drains the GPU).  The drivers follow InvCG2_a (lib/actions/ferm/invert/invcg2.cc:70-232)
// and InvBiCGStab_a (invbicgstab.cc:10-202) step by step; the shells follow
// LinOpSysSolverCG::operator() (syssolver_linop_cg.h:57-96) and LinOpSysSolverBiCGStab::operator()
// (syssolver_linop_bicgstab.h:57-95).
#pragma once
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <cmath>

#include "../../include/b200_clover.h"
#include "common.cuh"

struct b200_field {
  void* d;        // device planes C[nrhs][12][Vh]
  size_t bytes;
  int prec;
  int nrhs;       // right-hand sides held (1 for an ordinary field)
};

namespace b200 {

void set_error(const char* fmt, ...);

#define B200_CUDA(call)                                                                   \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess) {                                                              \
      b200::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      return B200_ERR_CUDA;                                                               \
    }                                                                                     \
  } while (0)

struct Config {
  int device;
  int gdims[4], pgrid[4], pcoord[4], ldims[4];
  int prec;
  b200_comm comm;
  bool have_comm;
};

// Precision-independent interface the C ABI talks to.
class EngineBase {
 public:
  virtual ~EngineBase() {}
  virtual int init() = 0;
  virtual int load_gauge(const void* const u[4], int host_prec, const double aniso[4], int t_boundary, int recon) = 0;
  virtual int load_clover(const void* clov, const void* invclov, int host_prec) = 0;
  virtual int make_clover(double diag_mass, double cr, double ct, int aniso, int t_dir) = 0;
  virtual int get_clover(void* clov, void* invclov, int host_prec) = 0;
  virtual int clover_logdet(double* out, int cb) = 0;
  virtual int set_preconditioning(int mode) = 0;
  virtual int set_twisted_mass(double mu) = 0;
  virtual int invert_multishift(b200_field* psi, const b200_field* chi, int n_shift, const double* shifts, const double* rsd,
                                int max_iter, b200_solve_info* info) = 0;   // info[n_shift]
  virtual int field_alloc(b200_field** f, int nrhs = 1) = 0;
  virtual void field_free(b200_field* f) = 0;
  virtual int field_upload(b200_field* f, const void* host, int host_prec, int irhs = 0) = 0;
  virtual int field_download(const b200_field* f, void* host, int host_prec, int irhs = 0) = 0;
  virtual int field_zero(b200_field* f) = 0;
  virtual int dslash(b200_field* out, const b200_field* in, int isign, int out_cb) = 0;
  virtual int clover_apply(b200_field* out, const b200_field* in, int cb, int inverse) = 0;
  virtual int matpc(b200_field* out, const b200_field* in, int isign) = 0;
  virtual int time_matpc(b200_field* out, const b200_field* in, int isign, int reps, double ms[2]) = 0;
  virtual int norm2(const b200_field* x, double* r) = 0;                          // r[nrhs]
  virtual int inner(const b200_field* x, const b200_field* y, double r[2]) = 0;   // r[nrhs][2]
  virtual int invert(b200_field* psi, const b200_field* chi, int solver, double rsd, int max_iter, int mdagm, b200_solve_info* info) = 0;   // info[nrhs]
  virtual int iterate_begin(b200_field* psi, const b200_field* chi, int solver) = 0;
  virtual int iterate(int solver, int n_iter) = 0;
  virtual int time_solver_kernels(int solver, int reps, double* ms, int max_ms, int* n_ms) = 0;
  virtual int qprop(void* psi, const void* chi, int host_prec, int nrhs, int solver, double rsd, int max_iter,
                    b200_solve_info* infos) = 0;
  virtual int sync() = 0;

  Config cfg;
  Geom g;
  cudaStream_t stream = nullptr;
  long long launches = 0;
  int host_copy_threads = 4;   // size of the host thread team of the pageable-copy pipeline (set by init)
};

// Mixed-precision reliable-update CG (engine_mixed.cu): hi must be the fp64 engine; *lo_slot is created on first use.
int reliable_solve(EngineBase* hi, EngineBase** lo_slot, b200_field* psi, const b200_field* chi, double rsd, double delta,
                   int max_iter, int mdagm, b200_solve_info* info);

// Mixed-precision reliable-update BiCGStab (engine_mixed.cu): RelInvBiCGStab_a, reliable_bicgstab.cc:13-290
int reliable_bicgstab_solve(EngineBase* hi, EngineBase** lo_slot, b200_field* psi, const b200_field* chi, double rsd, double delta,
                            int max_iter, int mdagm, b200_solve_info* info);

EngineBase* make_engine_double(const Config& c);
EngineBase* make_engine_float(const Config& c);

}  // namespace b200

struct b200_ctx {
  b200::EngineBase* eng;
  b200::EngineBase* sloppy;   // fp32 twin of an fp64 engine, made on demand by b200_invert_reliable
  // device twins of the host psi / chi of the host-pointer entry points: allocated on the first solve and kept for the
  // life of the context (Chroma calls operator() 12 times per propagator, quarkprop4_w.cc:70-117; a cudaMalloc/cudaFree
  // pair of 1 GB per call costs more than the transfers)
  b200_field* host_tmp[3];
  // device twin of the n_shift host solutions of b200_invert_multishift: kept (and grown on demand) for the life of the
  // context -- the rational monomials of RHMC call the multi-shift solver once per force / action evaluation
  b200_field* ms_sol;
};
