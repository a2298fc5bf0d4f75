// -*- C++ -*-
/*! \file
 *  \brief M psi = chi for the even-odd preconditioned clover operator on a B200 (plugin B200_CLOVER_INVERTER)
 *
 *  Twin of LinOpSysSolverQUDAClover (quda_solvers/syssolver_linop_clover_quda_w.h:71-648): same factory signature,
 *  same subset, same residual re-check with Chroma's own linop, same failure policy.
 */
#ifndef __SYSSOLVER_LINOP_CLOVER_B200_W_H__
#define __SYSSOLVER_LINOP_CLOVER_B200_W_H__

#include "chroma_config.h"

#ifdef BUILD_B200

#include "handle.h"
#include "state.h"
#include "syssolver.h"
#include "linearop.h"
#include "actions/ferm/invert/syssolver_linop.h"
#include "actions/ferm/invert/b200_solvers/syssolver_b200_clover_params.h"
#include "actions/ferm/invert/b200_solvers/b200_clover_engine.h"

namespace Chroma
{
  namespace LinOpSysSolverB200CloverEnv
  {
    bool registerAll();
  }

  class LinOpSysSolverB200Clover : public LinOpSystemSolver<LatticeFermion>
  {
  public:
    typedef LatticeFermion T;
    typedef LatticeColorMatrix U;
    typedef multi1d<LatticeColorMatrix> Q;

    LinOpSysSolverB200Clover(Handle< LinearOperator<T> > A_, Handle< FermState<T,Q,Q> > state_,
                             const SysSolverB200CloverParams& invParam_)
      : A(A_), invParam(invParam_), engine(new B200CloverEngine(state_, invParam_))
    {
      engine->checkOperator(*A);
      QDPIO::cout << "LinOpSysSolverB200Clover: engine ready" << std::endl;
    }

    ~LinOpSysSolverB200Clover() {}

    const Subset& subset() const { return A->subset(); }

    //! psi: initial guess in, solution out (rb[1]); chi: source (rb[1])
    SystemSolverResults_t operator()(T& psi, const T& chi) const
    {
      START_CODE();
      SystemSolverResults_t res;
      StopWatch swatch;
      swatch.start();

      const b200_solve_info info = engine->solve(psi, chi, false);
      res.n_count = info.n_count;

      // true residual with Chroma's operator, as the CPU and QUDA shells do
      {
        T r;
        r[A->subset()] = chi;
        T tmp;
        (*A)(tmp, psi, PLUS);
        r[A->subset()] -= tmp;
        res.resid = sqrt(norm2(r, A->subset()));
      }
      const Double rel_resid = res.resid / sqrt(norm2(chi, A->subset()));
      swatch.stop();
      QDPIO::cout << "B200_CLOVER_SOLVER: " << res.n_count << " iterations. Rsd = " << res.resid
                  << " Relative Rsd = " << rel_resid << "  (" << swatch.getTimeInSeconds() << " s)" << std::endl;

      if (!invParam.SilentFailP && toBool(rel_resid > invParam.RsdToleranceFactor * invParam.RsdTarget)) {
        QDPIO::cerr << "ERROR: B200 solver residuum is outside tolerance: resid=" << rel_resid
                    << " Desired=" << invParam.RsdTarget
                    << " Max Tolerated=" << invParam.RsdToleranceFactor * invParam.RsdTarget << std::endl;
        QDP_abort(1);
      }
      END_CODE();
      return res;
    }

  private:
    LinOpSysSolverB200Clover() {}

    Handle< LinearOperator<T> > A;
    const SysSolverB200CloverParams invParam;
    Handle< B200CloverEngine > engine;
  };
}

#endif // BUILD_B200
#endif
