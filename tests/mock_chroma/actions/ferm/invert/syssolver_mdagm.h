#ifndef MOCK_SYSSOLVER_MDAGM_H
#define MOCK_SYSSOLVER_MDAGM_H
#include "syssolver.h"
#include "update/molecdyn/predictor/chrono_predictor.h"
namespace Chroma {
template <typename T> class MdagMSystemSolver : public SystemSolver<T> {   // actions/ferm/invert/syssolver_mdagm.h
 public:
  virtual SystemSolverResults_t operator()(T& psi, const T& chi) const = 0;
  virtual SystemSolverResults_t operator()(T& psi, const T& chi, AbsChronologicalPredictor4D<T>& predictor) const = 0;
};
}
#endif
