#ifndef MOCK_LINEAROP_H
#define MOCK_LINEAROP_H
#include "chromabase.h"
namespace Chroma {
template <typename T> class LinearOperator {   // lib/linearop.h
 public:
  virtual ~LinearOperator() {}
  virtual void operator()(T& chi, const T& psi, enum PlusMinus isign) const = 0;
  virtual const Subset& subset() const = 0;
  virtual unsigned long nFlops() const { return 0; }
};
template <typename T, typename P, typename Q> class DiffLinearOperator : public LinearOperator<T> {};   // lib/linearop.h
}
#endif
