#!/bin/bash
# round-2 GPU job Q (1 GPU): batched kernel, neighbour-spinor prefetch depth 1 vs 2 (B200_MRHS_DEPTH) at NRB = 6 / 4 / 12
mkdir -p gpurun_out
for tag in "" d2n4 d2n12; do
  export B200_LIB_TAG=$tag
  echo "== variant '${tag:-product}'"
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_rhs or qprop or symmetric or twisted" > gpurun_out/r02q_pytest_${tag:-product}.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/r02q_pytest_${tag:-product}.log
  PROF_LATT=48,48,48,96 PROF_REPS=5 timeout 300 python scripts/prof_mrhs.py 2>&1 | tail -2
  timeout 300 python bench.py --no-cpu --no-fp32 --no-solve --steps 10 --warmup 3 > gpurun_out/r02q_bench_${tag:-product}.json 2> gpurun_out/r02q_bench_${tag:-product}.err
  python -c "
import json;b=json.loads(open('gpurun_out/r02q_bench_${tag:-product}.json').read().strip().splitlines()[-1])['multi_rhs'];print('fp64 M %.3f ms frac %.3f  cg iter %.2f ms'%(b['clover_dslash']['ms_per_apply'],b['clover_dslash']['frac_of_peak'],b['ms_per_iteration']))"
done
