#ifndef MOCK_LINOP_FACTORY_H
#define MOCK_LINOP_FACTORY_H
#include "factory_common.h"
#include "actions/ferm/invert/syssolver_linop.h"
namespace Chroma { typedef MockFactory< LinOpSystemSolver<LatticeFermion> > TheLinOpFermSystemSolverFactory; }
#endif
