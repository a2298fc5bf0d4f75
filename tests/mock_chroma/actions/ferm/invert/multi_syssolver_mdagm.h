#ifndef MOCK_MULTI_SYSSOLVER_MDAGM_H
#define MOCK_MULTI_SYSSOLVER_MDAGM_H
#include "syssolver.h"
namespace Chroma {
template <typename T> class MultiSystemSolver {   // lib/multi_syssolver.h
 public:
  virtual ~MultiSystemSolver() {}
  virtual SystemSolverResults_t operator()(multi1d<T>& psi, const multi1d<Real>& shifts, const T& chi) const = 0;
  virtual const Subset& subset() const = 0;
};
template <typename T> struct MdagMMultiSystemSolver : virtual public MultiSystemSolver<T> {};   // actions/ferm/invert/multi_syssolver_mdagm.h:18
}
#endif
