#ifndef MOCK_MDAGM_FACTORY_H
#define MOCK_MDAGM_FACTORY_H
#include "factory_common.h"
#include "actions/ferm/invert/syssolver_mdagm.h"
namespace Chroma { typedef MockFactory< MdagMSystemSolver<LatticeFermion> > TheMdagMFermSystemSolverFactory; }
#endif
