// -*- C++ -*-
/*! \file
 *  \brief RAII holder of a b200_ctx for one (gauge field, clover parameters) pair; shared by the LinOp and MdagM
 *         B200 system solvers.  All numerics happen behind the C ABI of include/b200_clover.h -- this file only
 *         hands QDP++ pointers across, exactly where the QUDA adapter hands them to loadGaugeQuda / loadCloverQuda /
 *         invertQuda (quda_solvers/syssolver_linop_clover_quda_w.h:514-552, syssolver_linop_clover_quda_w.cc:76-98).
 */
#ifndef __B200_CLOVER_ENGINE_H__
#define __B200_CLOVER_ENGINE_H__

#include "chromabase.h"
#include "handle.h"
#include "state.h"
#include "linearop.h"
#include "io/aniso_io.h"
#include "actions/ferm/invert/b200_solvers/syssolver_b200_clover_params.h"

#include <b200_clover.h>
#include <cstring>
#include <vector>

namespace Chroma
{
  namespace B200Glue
  {
    //! b200_comm callbacks on top of QDP++'s global sums (used ONCE, to swap CUDA IPC handles between the ranks)
    //! user -> int[2] = {rank, size} in the ENGINE's rank order (pt*Pz + pz), which need not be Layout::nodeNumber()
    inline int allgather(void* user, const void* send, void* recv, size_t bytes)
    {
      const int* rs = static_cast<const int*>(user);
      const int me = rs[0], nodes = rs[1];
      std::vector<int> buf(static_cast<size_t>(nodes) * bytes, 0);
      const unsigned char* s = static_cast<const unsigned char*>(send);
      for (size_t i = 0; i < bytes; ++i) buf[static_cast<size_t>(me) * bytes + i] = s[i];
      QDPInternal::globalSumArray(buf.data(), static_cast<int>(buf.size()));
      unsigned char* r = static_cast<unsigned char*>(recv);
      for (size_t i = 0; i < buf.size(); ++i) r[i] = static_cast<unsigned char>(buf[i]);
      return 0;
    }
    inline int barrier(void*)
    {
      int one = 1;
      QDPInternal::globalSum(one);
      return 0;
    }
  }

  class B200CloverEngine
  {
  public:
    typedef LatticeFermion T;
    typedef LatticeColorMatrix U;
    typedef multi1d<LatticeColorMatrix> Q;
    typedef WordType<T>::Type_t REALT;

    B200CloverEngine(Handle< FermState<T,Q,Q> > state, const SysSolverB200CloverParams& p) : ctx(0), invParam(p)
    {
      host_prec = (sizeof(REALT) == 4) ? B200_SINGLE : B200_DOUBLE;
      dev_prec = host_prec;
      if (p.precision == B200_PREC_SINGLE) dev_prec = B200_SINGLE;
      if (p.precision == B200_PREC_DOUBLE) dev_prec = B200_DOUBLE;
      const bool reliable = p.solverType == B200_RELIABLE_CG_SOLVER || p.solverType == B200_RELIABLE_BICGSTAB_SOLVER;
      mixed = (dev_prec == B200_DOUBLE) && (p.sloppyPrecision == B200_PREC_SINGLE || reliable);
      if (reliable && dev_prec != B200_DOUBLE) {
        QDPIO::cerr << "B200_CLOVER_INVERTER: RELIABLE_CG / RELIABLE_BICGSTAB need CudaPrecision DOUBLE" << std::endl;
        QDP_abort(1);
      }

      // geometry: the engine splits T, or T and Z (1 x 1 x Pz x Pt; one rank per GPU inside one NVSwitch box)
      int gdims[4], grid[4], coord[4];
      const multi1d<int>& machsize = Layout::logicalSize();
      const multi1d<int>& mycoord = Layout::nodeCoord();
      for (int mu = 0; mu < Nd; ++mu) {
        gdims[mu] = Layout::lattSize()[mu];
        grid[mu] = machsize[mu];
        coord[mu] = mycoord[mu];
      }
      b200_comm comm;
      comm_rs[0] = coord[3] * grid[2] + coord[2]; comm_rs[1] = grid[2] * grid[3];
      comm.rank = comm_rs[0]; comm.size = comm_rs[1];
      comm.allgather = B200Glue::allgather; comm.barrier = B200Glue::barrier; comm.user = comm_rs;
      int device = p.device;
      if (device < 0) {
        const int ndev = b200_device_count();
        if (ndev <= 0) fail("no CUDA device visible");
        device = Layout::nodeNumber() % ndev;
      }
      check(b200_create(&ctx, device, gdims, grid, coord, Layout::numNodes() > 1 ? &comm : 0, dev_prec), "b200_create");

      // links: state->getLinks() already carry the fermion BC phases (simple_fermbc.h:87-103); they go over UNSCALED,
      // the anisotropy factors of the hopping term travel separately (makeFermCoeffs, io/aniso_io.cc:63-80)
      const Q& links = state->getLinks();
      const void* gauge[4];
      for (int mu = 0; mu < Nd; ++mu)
        gauge[mu] = (const void*)&(links[mu].elem(all.start()).elem().elem(0,0).real());
      double cf[4] = {1.0, 1.0, 1.0, 1.0};
      const AnisoParam_t& aniso = p.CloverParams.anisoParam;
      if (aniso.anisoP) {
        multi1d<Real> c = makeFermCoeffs(aniso);
        for (int mu = 0; mu < Nd; ++mu) cf[mu] = toDouble(c[mu]);
      }
      check(b200_load_gauge(ctx, gauge, host_prec, cf, p.AntiPeriodicT ? -1 : +1,
                            p.reconstruct == B200_RECONS_12_T ? B200_RECONS_12 : B200_RECONS_NONE), "b200_load_gauge");

      // clover term and its even-even inverse are built on the GPU from those links with the coefficients
      // QDPCloverTermT::create derives (clover_term_qdp_w.h:263-278) -- no host-side CloverTerm is needed
      double ff = 1.0, fm = 1.0;
      if (aniso.anisoP) { ff = 1.0 / toDouble(aniso.xi_0); fm = toDouble(aniso.nu) / toDouble(aniso.xi_0); }
      const double diag_mass = 1.0 + (Nd - 1) * fm + toDouble(p.CloverParams.Mass);
      const double clov_r = 0.5 * ff * toDouble(p.CloverParams.clovCoeffR);
      const double clov_t = 0.5 * toDouble(p.CloverParams.clovCoeffT);
      check(b200_make_clover(ctx, diag_mass, clov_r, clov_t, aniso.anisoP ? 1 : 0, aniso.t_dir), "b200_make_clover");
      // SymEvenOddPrecCloverLinOp::create also inverts the odd block (seoprec_clover_linop_w.cc:31-33): done on the GPU
      if (p.SymmetricLinopP) check(b200_set_preconditioning(ctx, B200_PRECOND_SYMMETRIC), "b200_set_preconditioning");
      // twisted-mass term of the clover operators (eoprec_clover_linop_w.cc:174-184, seoprec_clover_linop_w.cc:174-184)
      if (p.CloverParams.twisted_m_usedP)
        check(b200_set_twisted_mass(ctx, toDouble(p.CloverParams.twisted_m)), "b200_set_twisted_mass");
    }

    ~B200CloverEngine() { if (ctx) b200_destroy(ctx); }

    //! Guard against a parameter group that does not describe the operator the factory handed us (the creator's
    //! signature carries no such information): apply the caller's A and the engine's M to one test vector and
    //! compare.  Catches a wrong SymmetricLinop, Mass / clovCoeff / anisotropy / TwistedM that differ from the fermion
    //! action's, or a wrong AntiPeriodicT, at construction instead of as a failed residual check after the first solve.
    //! The test vector is Gaussian but QDP++'s global RNG is left exactly as found (RNG::savern / setrn): solvers are
    //! constructed many times per HMC trajectory and a run must reproduce the random sequence of the same XML and seed
    //! run with the CPU or QUDA solvers (their constructors draw nothing).  <CheckOperator>false</CheckOperator> skips it.
    void checkOperator(const LinearOperator<T>& A) const
    {
      if (!invParam.CheckOperatorP) return;
      T x = zero, want = zero, got = zero;
      Seed saved;
      QDP::RNG::savern(saved);
      gaussian(x, rb[1]);
      QDP::RNG::setrn(saved);
      A(want, x, PLUS);
      const void* in = (const void*)&(x.elem(rb[1].start()).elem(0).elem(0).real());
      void* out = (void*)&(got.elem(rb[1].start()).elem(0).elem(0).real());
      check(b200_clover_matpc(ctx, out, in, host_prec, B200_PLUS), "b200_clover_matpc");
      got[rb[1]] -= want;
      const Double rel = sqrt(norm2(got, rb[1])) / sqrt(norm2(want, rb[1]));
      const Double tol = (dev_prec == B200_DOUBLE && host_prec == B200_DOUBLE) ? Double(1.0e-10) : Double(1.0e-4);
      if (toBool(rel > tol)) {
        QDPIO::cerr << "B200_CLOVER_INVERTER: the engine's operator differs from the fermion action's linear operator (rel. diff "
                    << rel << "): check SymmetricLinop, CloverParams and AntiPeriodicT in the InvertParam group" << std::endl;
        QDP_abort(1);
      }
    }

    //! psi, chi live on rb[1]; cb2 layout makes that one contiguous block starting at rb[1].start()
    b200_solve_info solve(T& psi, const T& chi, bool mdagm) const
    {
      void* out = (void*)&(psi.elem(rb[1].start()).elem(0).elem(0).real());
      const void* in = (const void*)&(chi.elem(rb[1].start()).elem(0).elem(0).real());
      b200_solve_info info;
      std::memset(&info, 0, sizeof(info));
      const double rsd = toDouble(invParam.RsdTarget);
      int rc;
      const bool bicg = invParam.solverType == B200_BICGSTAB_SOLVER || invParam.solverType == B200_RELIABLE_BICGSTAB_SOLVER;
      if (mixed && !bicg)
        rc = b200_invert_reliable(ctx, out, in, host_prec, rsd, toDouble(invParam.Delta), invParam.MaxIter, mdagm ? 1 : 0, &info);
      else if (mixed)
        rc = b200_invert_reliable_bicgstab(ctx, out, in, host_prec, rsd, toDouble(invParam.Delta), invParam.MaxIter, mdagm ? 1 : 0, &info);
      else {
        const int solver = bicg ? B200_SOLVER_BICGSTAB : B200_SOLVER_CG;
        rc = mdagm ? b200_invert_mdagm(ctx, out, in, host_prec, solver, rsd, invParam.MaxIter, &info)
                   : b200_invert(ctx, out, in, host_prec, solver, rsd, invParam.MaxIter, &info);
      }
      check(rc, "b200_invert");
      if (invParam.verboseP)
        QDPIO::cout << "B200_CLOVER_SOLVER: time=" << info.secs << " s\tPerformance=" << info.gflops
                    << " GFLOPS\tTotal Time (incl. copies)=" << info.secs_total << " s" << std::endl;
      return info;
    }

    //! (M^dag M + shifts[s]) psi[s] = chi on rb[1]; psi[s] are zeroed by the engine like MInvCG2_a does (minvcg2.cc:118-122)
    std::vector<b200_solve_info> solveMulti(multi1d<T>& psi, const multi1d<Real>& shifts, const T& chi) const
    {
      const int n = shifts.size();
      if (n < 1 || n > B200_MAX_SHIFTS) fail("multi-shift solve: number of shifts out of range");
      if (psi.size() < n) psi.resize(n);
      std::vector<void*> out(n);
      std::vector<double> sh(n), rsd(n, toDouble(invParam.RsdTarget));
      for (int s = 0; s < n; ++s) {
        psi[s][rb[1]] = zero;     // touch the field so that QDP++ has allocated it before we take its address
        out[s] = (void*)&(psi[s].elem(rb[1].start()).elem(0).elem(0).real());
        sh[s] = toDouble(shifts[s]);
      }
      const void* in = (const void*)&(chi.elem(rb[1].start()).elem(0).elem(0).real());
      std::vector<b200_solve_info> info(n);
      std::memset(info.data(), 0, sizeof(b200_solve_info) * n);
      check(b200_invert_multishift(ctx, out.data(), in, host_prec, n, sh.data(), rsd.data(), invParam.MaxIter, info.data()),
            "b200_invert_multishift");
      if (invParam.verboseP)
        QDPIO::cout << "B200_CLOVER_SOLVER (multi-shift): time=" << info[0].secs << " s\tPerformance=" << info[0].gflops
                    << " GFLOPS\tTotal Time (incl. copies)=" << info[0].secs_total << " s" << std::endl;
      return info;
    }

  private:
    void fail(const char* what) const
    {
      QDPIO::cerr << "B200_CLOVER_INVERTER: " << what << std::endl;
      QDP_abort(1);
    }
    void check(int rc, const char* call) const
    {
      if (rc != B200_OK) {
        QDPIO::cerr << "B200_CLOVER_INVERTER: " << call << " failed (" << rc << "): " << b200_last_error() << std::endl;
        QDP_abort(1);
      }
    }

    b200_ctx* ctx;
    int comm_rs[2];   // {rank, size} handed to the b200_comm callbacks
    const SysSolverB200CloverParams invParam;
    int host_prec;
    int dev_prec;
    bool mixed;
  };
}

#endif
