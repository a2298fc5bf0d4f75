// -*- C++ -*-
/*! \file
 *  \brief Parameters of the B200 clover system solver (XML group <invType>B200_CLOVER_INVERTER</invType>)
 *
 *  Lives in lib/actions/ferm/invert/b200_solvers/ of a Chroma tree built with --with-b200=<dir>.
 *  The schema is the subset of SysSolverQUDACloverParams (quda_solvers/syssolver_quda_clover_params.h) that has a
 *  meaning for this engine, under the same tag names where one exists, so that an existing QUDA XML group works after
 *  changing <invType>: MaxIter, RsdTarget, CloverParams, AntiPeriodicT, SolverType, Delta, CudaPrecision,
 *  CudaSloppyPrecision, CudaReconstruct, RsdToleranceFactor, SilentFail, Verbose; own tags: Device, SymmetricLinop, CheckOperator.  Tags this engine ignores
 *  (AsymmetricLinop -- QUDA's internal choice; here the operator solved is always the caller's own A, see
 *  SymmetricLinop --, AxialGaugeFix, AutotuneDslash, Pipeline,
 *  GCRInnerParams, BackupSolverParam, DumpOnFail) are accepted and skipped.
 */
#ifndef __SYSSOLVER_B200_CLOVER_PARAMS_H__
#define __SYSSOLVER_B200_CLOVER_PARAMS_H__

#include "chromabase.h"
#include "actions/ferm/fermacts/clover_fermact_params_w.h"
#include <string>

namespace Chroma
{
  enum B200SolverType { B200_CG_SOLVER, B200_BICGSTAB_SOLVER, B200_RELIABLE_CG_SOLVER, B200_RELIABLE_BICGSTAB_SOLVER };
  enum B200PrecisionType { B200_PREC_DEFAULT, B200_PREC_SINGLE, B200_PREC_DOUBLE };
  enum B200ReconsType { B200_RECONS_NONE_T, B200_RECONS_12_T };

  struct SysSolverB200CloverParams
  {
    SysSolverB200CloverParams(XMLReader& xml, const std::string& path);
    SysSolverB200CloverParams();

    CloverFermActParams CloverParams;
    bool AntiPeriodicT;
    int MaxIter;
    Real RsdTarget;
    Real Delta;                        //!< reliable-update threshold (RELIABLE_CG / RELIABLE_BICGSTAB, or CG / BICGSTAB with a sloppy precision)
    B200SolverType solverType;
    B200PrecisionType precision;       //!< device precision (DEFAULT = precision of the Chroma build)
    B200PrecisionType sloppyPrecision; //!< SINGLE below a DOUBLE precision selects the mixed-precision solver
    B200ReconsType reconstruct;        //!< default RECONS_NONE: bit-level parity with Chroma's own Dslash
    bool SilentFailP;
    Real RsdToleranceFactor;
    bool verboseP;
    int device;                        //!< CUDA device; -1 = node number modulo visible devices
    //! false (default): the fermion action hands the plugin an EvenOddPrecCloverLinOp (CLOVER, clover_fermact_w.cc);
    //! true: a SymEvenOddPrecCloverLinOp (SEOPREC_CLOVER, seoprec_clover_fermact_w.cc).  The engine must solve the
    //! caller's own A for the plugin's residual check with A to pass.
    bool SymmetricLinopP;
    //! true (default): compare the engine's operator with the caller's A on one vector at construction (RNG state preserved)
    bool CheckOperatorP;
  };

  void read(XMLReader& xml, const std::string& path, SysSolverB200CloverParams& p);
  void write(XMLWriter& xml, const std::string& path, const SysSolverB200CloverParams& p);
}

#endif
