#ifndef MOCK_ANISO_IO_H
#define MOCK_ANISO_IO_H
#include "chromabase.h"
namespace Chroma {
struct AnisoParam_t {   // lib/io/aniso_io.h
  AnisoParam_t() : anisoP(false), t_dir(3), xi_0(1), nu(1) {}
  bool anisoP; int t_dir; Real xi_0; Real nu;
};
multi1d<Real> makeFermCoeffs(const AnisoParam_t& aniso);                 // lib/io/aniso_io.cc:63-80
}
#endif
