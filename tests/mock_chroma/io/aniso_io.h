#ifndef MOCK_ANISO_IO_H
#define MOCK_ANISO_IO_H
#include "chromabase.h"
namespace Chroma {
struct AnisoParam_t { bool anisoP; int t_dir; Real xi_0; Real nu; };   // lib/io/aniso_io.h
multi1d<Real> makeFermCoeffs(const AnisoParam_t& aniso);                 // lib/io/aniso_io.cc:63-80
}
#endif
