#ifndef MOCK_SYSSOLVER_H
#define MOCK_SYSSOLVER_H
#include "chromabase.h"
namespace Chroma {
struct SystemSolverResults_t { SystemSolverResults_t() : n_count(0), resid(0) {} int n_count; Real resid; };   // lib/syssolver.h:16-23
template <typename T> class SystemSolver {
 public:
  virtual ~SystemSolver() {}
  virtual SystemSolverResults_t operator()(T& psi, const T& chi) const = 0;
  virtual const Subset& subset() const = 0;
};
}
#endif
