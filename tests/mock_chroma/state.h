#ifndef MOCK_STATE_H
#define MOCK_STATE_H
#include "chromabase.h"
namespace Chroma {
template <typename T, typename P, typename Q> class FermState {   // lib/state.h
 public:
  virtual ~FermState() {}
  virtual const Q& getLinks() const = 0;
};
}
#endif
