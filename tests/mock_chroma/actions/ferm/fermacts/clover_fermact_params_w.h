#ifndef MOCK_CLOVER_PARAMS_H
#define MOCK_CLOVER_PARAMS_H
#include "io/aniso_io.h"
namespace Chroma {
struct CloverFermActParams {   // lib/actions/ferm/fermacts/clover_fermact_params_w.h:17-37
  CloverFermActParams() : Mass(0), clovCoeffR(0), clovCoeffT(0), u0(1), twisted_m(0), twisted_m_usedP(false) {}
  Real Mass, clovCoeffR, clovCoeffT, u0;
  AnisoParam_t anisoParam;
  Real twisted_m;
  bool twisted_m_usedP;
};
void read(XMLReader&, const std::string&, CloverFermActParams&);
void write(XMLWriter&, const std::string&, const CloverFermActParams&);
}
#endif
