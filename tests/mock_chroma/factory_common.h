#ifndef MOCK_FACTORY_COMMON_H
#define MOCK_FACTORY_COMMON_H
#include <map>
#include "chromabase.h"
#include "handle.h"
#include "state.h"
#include "linearop.h"
namespace Chroma {
// lib/objfactory.h + singleton.h with the 4-argument creator of the fermion system-solver factories
// (syssolver_linop_factory.h:29-39): registerObject / createObject by name.
template <typename Product> class MockFactory {
 public:
  typedef Product* (*Creator)(XMLReader&, const std::string&,
                              Handle< FermState< LatticeFermion, multi1d<LatticeColorMatrix>, multi1d<LatticeColorMatrix> > >,
                              Handle< LinearOperator<LatticeFermion> >);
  static MockFactory& Instance() { static MockFactory f; return f; }
  bool registerObject(const std::string& name, Creator c) { return creators.insert(std::make_pair(name, c)).second; }
  Product* createObject(const std::string& name, XMLReader& xml, const std::string& path,
                        Handle< FermState< LatticeFermion, multi1d<LatticeColorMatrix>, multi1d<LatticeColorMatrix> > > state,
                        Handle< LinearOperator<LatticeFermion> > A) {
    typename std::map<std::string, Creator>::const_iterator it = creators.find(name);
    if (it == creators.end()) { QDPIO::cerr << "factory: unknown object " << name << std::endl; QDP_abort(1); }
    return (it->second)(xml, path, state, A);
  }
 private:
  std::map<std::string, Creator> creators;
};
}
#endif
