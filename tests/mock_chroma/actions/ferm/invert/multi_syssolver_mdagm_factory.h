#ifndef MOCK_MULTI_MDAGM_FACTORY_H
#define MOCK_MULTI_MDAGM_FACTORY_H
#include "factory_common.h"
#include "actions/ferm/invert/multi_syssolver_mdagm.h"
namespace Chroma { typedef MockFactory< MdagMMultiSystemSolver<LatticeFermion> > TheMdagMFermMultiSystemSolverFactory; }
#endif
