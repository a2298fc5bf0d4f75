#ifndef MOCK_FACTORY_COMMON_H
#define MOCK_FACTORY_COMMON_H
#include "chromabase.h"
#include "handle.h"
#include "state.h"
#include "linearop.h"
namespace Chroma {
template <typename Product> class MockFactory {   // lib/objfactory.h + singleton.h, 4-argument creator
 public:
  typedef Product* (*Creator)(XMLReader&, const std::string&,
                              Handle< FermState< LatticeFermion, multi1d<LatticeColorMatrix>, multi1d<LatticeColorMatrix> > >,
                              Handle< LinearOperator<LatticeFermion> >);
  static MockFactory& Instance() { static MockFactory f; return f; }
  bool registerObject(const std::string&, Creator) { return true; }
};
}
#endif
