// -*- C++ -*-
/*! \file
 *  \brief M^dag M psi = chi for the even-odd preconditioned clover operator on a B200 (HMC call sites)
 *
 *  Twin of MdagMSysSolverQUDAClover (quda_solvers/syssolver_mdagm_clover_quda_w.h): registered in
 *  TheMdagMFermSystemSolverFactory (syssolver_mdagm_factory.h:30-39) under the same key B200_CLOVER_INVERTER, used by
 *  TwoFlavorExactWilsonTypeFermMonomial (update/molecdyn/monomial/two_flavor_monomial_w.h:74-87).
 */
#ifndef __SYSSOLVER_MDAGM_CLOVER_B200_W_H__
#define __SYSSOLVER_MDAGM_CLOVER_B200_W_H__

#include "chroma_config.h"

#ifdef BUILD_B200

#include "handle.h"
#include "state.h"
#include "syssolver.h"
#include "linearop.h"
#include "lmdagm.h"
#include "actions/ferm/invert/syssolver_mdagm.h"
#include "update/molecdyn/predictor/chrono_predictor.h"
#include "actions/ferm/invert/b200_solvers/syssolver_b200_clover_params.h"
#include "actions/ferm/invert/b200_solvers/b200_clover_engine.h"

namespace Chroma
{
  namespace MdagMSysSolverB200CloverEnv
  {
    bool registerAll();
  }

  class MdagMSysSolverB200Clover : public MdagMSystemSolver<LatticeFermion>
  {
  public:
    typedef LatticeFermion T;
    typedef LatticeColorMatrix U;
    typedef multi1d<LatticeColorMatrix> Q;

    MdagMSysSolverB200Clover(Handle< LinearOperator<T> > A_, Handle< FermState<T,Q,Q> > state_,
                             const SysSolverB200CloverParams& invParam_)
      : A(A_), invParam(invParam_), engine(new B200CloverEngine(state_, invParam_))
    {
      engine->checkOperator(*A);
    }

    ~MdagMSysSolverB200Clover() {}

    const Subset& subset() const { return A->subset(); }

    //! psi: initial guess in, solution out
    SystemSolverResults_t operator()(T& psi, const T& chi) const
    {
      START_CODE();
      SystemSolverResults_t res;
      const b200_solve_info info = engine->solve(psi, chi, true);
      res.n_count = info.n_count;
      {
        T tmp = zero, r = zero;
        (*A)(tmp, psi, PLUS);
        (*A)(r, tmp, MINUS);
        r[A->subset()] -= chi;
        res.resid = sqrt(norm2(r, A->subset()));
      }
      const Double rel_resid = res.resid / sqrt(norm2(chi, A->subset()));
      QDPIO::cout << "B200_CLOVER_SOLVER (MdagM): " << res.n_count << " iterations. Rsd = " << res.resid
                  << " Relative Rsd = " << rel_resid << std::endl;
      if (!invParam.SilentFailP && toBool(rel_resid > invParam.RsdToleranceFactor * invParam.RsdTarget)) {
        QDPIO::cerr << "ERROR: B200 MdagM solver residuum is outside tolerance: resid=" << rel_resid
                    << " Desired=" << invParam.RsdTarget << std::endl;
        QDP_abort(1);
      }
      END_CODE();
      return res;
    }

    //! with a chronological guess (same contract as MdagMSysSolverCG, syssolver_mdagm_cg.h:104-123)
    SystemSolverResults_t operator()(T& psi, const T& chi, AbsChronologicalPredictor4D<T>& predictor) const
    {
      START_CODE();
      {
        Handle< LinearOperator<T> > MdagM(new MdagMLinOp<T>(A));
        predictor(psi, (*MdagM), chi);
      }
      SystemSolverResults_t res = (*this)(psi, chi);
      predictor.newVector(psi);
      END_CODE();
      return res;
    }

  private:
    MdagMSysSolverB200Clover() {}

    Handle< LinearOperator<T> > A;
    const SysSolverB200CloverParams invParam;
    Handle< B200CloverEngine > engine;
  };
}

#endif // BUILD_B200
#endif
