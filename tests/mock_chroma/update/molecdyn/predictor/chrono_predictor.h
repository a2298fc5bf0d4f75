#ifndef MOCK_CHRONO_H
#define MOCK_CHRONO_H
#include "linearop.h"
namespace Chroma {
template <typename T> class AbsChronologicalPredictor4D {   // chrono_predictor.h:23-52
 public:
  virtual ~AbsChronologicalPredictor4D() {}
  virtual void operator()(T& psi, const LinearOperator<T>& A, const T& chi) = 0;
  virtual void reset() = 0;
  virtual void newVector(const T& psi) = 0;
};
}
#endif
