#ifndef MOCK_CLOVER_PARAMS_H
#define MOCK_CLOVER_PARAMS_H
#include "io/aniso_io.h"
namespace Chroma {
struct CloverFermActParams {   // clover_fermact_params_w.h
  Real Mass, clovCoeffR, clovCoeffT, u0;
  AnisoParam_t anisoParam;
};
void read(XMLReader&, const std::string&, CloverFermActParams&);
void write(XMLWriter&, const std::string&, const CloverFermActParams&);
}
#endif
