/*! \file
 *  \brief Registration of B200_CLOVER_INVERTER in TheMdagMFermMultiSystemSolverFactory
 */
#include "chroma_config.h"

#ifdef BUILD_B200

#include "actions/ferm/invert/multi_syssolver_mdagm_factory.h"
#include "actions/ferm/invert/multi_syssolver_mdagm_aggregate.h"
#include "actions/ferm/invert/b200_solvers/multi_syssolver_mdagm_clover_b200_w.h"

namespace Chroma
{
  namespace MdagMMultiSysSolverB200CloverEnv
  {
    namespace
    {
      const std::string name("B200_CLOVER_INVERTER");
      bool registered = false;
    }

    MdagMMultiSystemSolver<LatticeFermion>* createFerm(XMLReader& xml_in,
                                                       const std::string& path,
                                                       Handle< FermState< LatticeFermion, multi1d<LatticeColorMatrix>, multi1d<LatticeColorMatrix> > > state,
                                                       Handle< LinearOperator<LatticeFermion> > A)
    {
      return new MdagMMultiSysSolverB200Clover(A, state, SysSolverB200CloverParams(xml_in, path));
    }

    bool registerAll()
    {
      bool success = true;
      if (!registered) {
        success &= Chroma::TheMdagMFermMultiSystemSolverFactory::Instance().registerObject(name, createFerm);
        registered = true;
      }
      return success;
    }
  }
}

#endif
