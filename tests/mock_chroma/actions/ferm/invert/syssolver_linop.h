#ifndef MOCK_SYSSOLVER_LINOP_H
#define MOCK_SYSSOLVER_LINOP_H
#include "syssolver.h"
namespace Chroma { template <typename T> class LinOpSystemSolver : public SystemSolver<T> {}; }
#endif
