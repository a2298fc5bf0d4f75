#ifndef MOCK_LMDAGM_H
#define MOCK_LMDAGM_H
#include "linearop.h"
#include "handle.h"
namespace Chroma {
template <typename T> class MdagMLinOp : public LinearOperator<T> {   // lib/actions/ferm/linop/lmdagm.h: chi = A^dag A psi
 public:
  MdagMLinOp(Handle< LinearOperator<T> > A_) : A(A_) {}
  void operator()(T& chi, const T& psi, enum PlusMinus) const { T tmp = zero; (*A)(tmp, psi, PLUS); (*A)(chi, tmp, MINUS); }
  const Subset& subset() const { return A->subset(); }
 private:
  Handle< LinearOperator<T> > A;
};
}
#endif
