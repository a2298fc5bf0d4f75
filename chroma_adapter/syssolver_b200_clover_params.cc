/*! \file
 *  \brief Reader/writer of the B200_CLOVER_INVERTER parameter group
 */
#include "actions/ferm/invert/b200_solvers/syssolver_b200_clover_params.h"

using namespace QDP;

namespace Chroma
{
  namespace
  {
    template <typename E>
    bool lookup(const std::string& s, const char* const* names, int n, E& out)
    {
      for (int i = 0; i < n; ++i)
        if (s == names[i]) { out = static_cast<E>(i); return true; }
      return false;
    }
    const char* const solverNames[] = {"CG", "BICGSTAB", "RELIABLE_CG", "RELIABLE_BICGSTAB"};
    const char* const precNames[] = {"DEFAULT", "SINGLE", "DOUBLE"};
    const char* const reconsNames[] = {"RECONS_NONE", "RECONS_12"};

    template <typename E>
    void readEnum(XMLReader& top, const std::string& tag, const char* const* names, int n, E& out, E dflt, bool required)
    {
      if (top.count(tag) == 0) {
        if (required) { QDPIO::cerr << "B200_CLOVER_INVERTER: missing tag " << tag << std::endl; QDP_abort(1); }
        out = dflt;
        return;
      }
      std::string s;
      read(top, tag, s);
      if (!lookup(s, names, n, out)) {
        QDPIO::cerr << "B200_CLOVER_INVERTER: unknown value '" << s << "' for " << tag << std::endl;
        QDP_abort(1);
      }
    }
  }

  SysSolverB200CloverParams::SysSolverB200CloverParams()
    : AntiPeriodicT(true), MaxIter(5000), RsdTarget(Real(1.0e-8)), Delta(Real(0.1)), solverType(B200_CG_SOLVER),
      precision(B200_PREC_DEFAULT), sloppyPrecision(B200_PREC_DEFAULT), reconstruct(B200_RECONS_NONE_T),
      SilentFailP(false), RsdToleranceFactor(Real(10)), verboseP(false), device(-1), SymmetricLinopP(false), CheckOperatorP(true)
  {}

  SysSolverB200CloverParams::SysSolverB200CloverParams(XMLReader& xml, const std::string& path)
  {
    *this = SysSolverB200CloverParams();
    XMLReader paramtop(xml, path);

    read(paramtop, "MaxIter", MaxIter);
    read(paramtop, "RsdTarget", RsdTarget);
    read(paramtop, "CloverParams", CloverParams);
    read(paramtop, "AntiPeriodicT", AntiPeriodicT);
    readEnum(paramtop, "SolverType", solverNames, 4, solverType, B200_CG_SOLVER, true);

    if (paramtop.count("Delta") > 0) read(paramtop, "Delta", Delta);
    readEnum(paramtop, "CudaPrecision", precNames, 3, precision, B200_PREC_DEFAULT, false);
    readEnum(paramtop, "CudaSloppyPrecision", precNames, 3, sloppyPrecision, B200_PREC_DEFAULT, false);
    readEnum(paramtop, "CudaReconstruct", reconsNames, 2, reconstruct, B200_RECONS_NONE_T, false);
    if (paramtop.count("SilentFail") > 0) read(paramtop, "SilentFail", SilentFailP);
    if (paramtop.count("RsdToleranceFactor") > 0) read(paramtop, "RsdToleranceFactor", RsdToleranceFactor);
    if (paramtop.count("Verbose") > 0) read(paramtop, "Verbose", verboseP);
    if (paramtop.count("Device") > 0) read(paramtop, "Device", device);
    if (paramtop.count("SymmetricLinop") > 0) read(paramtop, "SymmetricLinop", SymmetricLinopP);
    if (paramtop.count("CheckOperator") > 0) read(paramtop, "CheckOperator", CheckOperatorP);
  }

  void read(XMLReader& xml, const std::string& path, SysSolverB200CloverParams& p)
  {
    SysSolverB200CloverParams tmp(xml, path);
    p = tmp;
  }

  void write(XMLWriter& xml, const std::string& path, const SysSolverB200CloverParams& p)
  {
    push(xml, path);
    write(xml, "invType", std::string("B200_CLOVER_INVERTER"));
    write(xml, "MaxIter", p.MaxIter);
    write(xml, "RsdTarget", p.RsdTarget);
    write(xml, "CloverParams", p.CloverParams);
    write(xml, "AntiPeriodicT", p.AntiPeriodicT);
    write(xml, "SolverType", std::string(solverNames[p.solverType]));
    write(xml, "Delta", p.Delta);
    write(xml, "CudaPrecision", std::string(precNames[p.precision]));
    write(xml, "CudaSloppyPrecision", std::string(precNames[p.sloppyPrecision]));
    write(xml, "CudaReconstruct", std::string(reconsNames[p.reconstruct]));
    write(xml, "SilentFail", p.SilentFailP);
    write(xml, "RsdToleranceFactor", p.RsdToleranceFactor);
    write(xml, "Verbose", p.verboseP);
    write(xml, "Device", p.device);
    write(xml, "SymmetricLinop", p.SymmetricLinopP);
    write(xml, "CheckOperator", p.CheckOperatorP);
    pop(xml);
  }
}
