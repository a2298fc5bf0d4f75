#!/bin/bash
# round-2 GPU job AE (1 GPU): quick confirmation of the last rebuild (host-side lattice-size guard only): smoke + a parity subset
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02ae_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02ae_smoke.log
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "matpc_parity or multi_rhs_operator or solver_matches or split_reduction" > gpurun_out/r02ae_pytest.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/r02ae_pytest.log
