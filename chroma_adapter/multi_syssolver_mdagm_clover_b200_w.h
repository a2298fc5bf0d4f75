// -*- C++ -*-
/*! \file
 *  \brief (M^dag M + shift_i) psi_i = chi for the even-odd preconditioned clover operator on a B200 (RHMC call sites)
 *
 *  Twin of MdagMMultiSysSolverCGQudaClover (quda_solvers/multi_syssolver_mdagm_cg_clover_quda_w.h) and drop-in for
 *  MdagMMultiSysSolverCG (multi_syssolver_mdagm_cg.h:31-118): registered in TheMdagMFermMultiSystemSolverFactory
 *  (multi_syssolver_mdagm_factory.h) under the key B200_CLOVER_INVERTER; used by the rational monomials
 *  (update/molecdyn/monomial/one_flavor_rat_monomial_w.h) through FermAct::mInvMdagM.
 */
#ifndef __MULTI_SYSSOLVER_MDAGM_CLOVER_B200_W_H__
#define __MULTI_SYSSOLVER_MDAGM_CLOVER_B200_W_H__

#include "chroma_config.h"

#ifdef BUILD_B200

#include "handle.h"
#include "state.h"
#include "syssolver.h"
#include "linearop.h"
#include "actions/ferm/invert/multi_syssolver_mdagm.h"
#include "actions/ferm/invert/b200_solvers/syssolver_b200_clover_params.h"
#include "actions/ferm/invert/b200_solvers/b200_clover_engine.h"

namespace Chroma
{
  namespace MdagMMultiSysSolverB200CloverEnv
  {
    bool registerAll();
  }

  class MdagMMultiSysSolverB200Clover : public MdagMMultiSystemSolver<LatticeFermion>
  {
  public:
    typedef LatticeFermion T;
    typedef LatticeColorMatrix U;
    typedef multi1d<LatticeColorMatrix> Q;

    MdagMMultiSysSolverB200Clover(Handle< LinearOperator<T> > A_, Handle< FermState<T,Q,Q> > state_,
                                  const SysSolverB200CloverParams& invParam_)
      : A(A_), invParam(invParam_), engine(new B200CloverEngine(state_, invParam_))
    {
      engine->checkOperator(*A);
    }

    ~MdagMMultiSysSolverB200Clover() {}

    const Subset& subset() const { return A->subset(); }

    //! psi[i] out (zero initial guesses, as MInvCG2 requires); RsdTarget applies to every shift
    SystemSolverResults_t operator()(multi1d<T>& psi, const multi1d<Real>& shifts, const T& chi) const
    {
      START_CODE();
      SystemSolverResults_t res;
      const std::vector<b200_solve_info> info = engine->solveMulti(psi, shifts, chi);
      res.n_count = info[0].n_count;
      if (!info[0].converged && !invParam.SilentFailP) {
        QDPIO::cerr << "B200_CLOVER_SOLVER (multi-shift): too many CG iterations: " << res.n_count << std::endl;   // minvcg2.cc:365-367
        QDP_abort(1);
      }
      // re-check every shift with Chroma's own A (the check the QUDA twin does under CheckShifts,
      // multi_syssolver_mdagm_cg_clover_quda_w.h, and MdagMMultiSysSolverCG logs, multi_syssolver_mdagm_cg.h:84-99)
      const Double chinorm = sqrt(norm2(chi, A->subset()));
      for (int i = 0; i < shifts.size(); ++i) {
        T tmp1 = zero, tmp2 = zero, r = zero;
        (*A)(tmp1, psi[i], PLUS);
        (*A)(tmp2, tmp1, MINUS);
        r[A->subset()] = chi;
        r[A->subset()] -= tmp2;
        T sp = zero;
        sp[A->subset()] = psi[i];
        sp[A->subset()] *= shifts[i];
        r[A->subset()] -= sp;
        const Double rel = sqrt(norm2(r, A->subset())) / chinorm;
        if (invParam.verboseP)
          QDPIO::cout << "B200_CLOVER_SOLVER (multi-shift): shift[" << i << "] = " << shifts[i] << " Relative Rsd = " << rel << std::endl;
        if (!invParam.SilentFailP && toBool(rel > invParam.RsdToleranceFactor * invParam.RsdTarget)) {
          QDPIO::cerr << "ERROR: B200 multi-shift solver residuum is outside tolerance: shift=" << i << " resid=" << rel
                      << " Desired=" << invParam.RsdTarget << std::endl;
          QDP_abort(1);
        }
        res.resid = rel * chinorm;
      }
      END_CODE();
      return res;
    }

  private:
    MdagMMultiSysSolverB200Clover() {}

    Handle< LinearOperator<T> > A;
    const SysSolverB200CloverParams invParam;
    Handle< B200CloverEngine > engine;
  };
}

#endif // BUILD_B200
#endif
