#!/bin/bash
# round-2 GPU job S (1 GPU): batched kernel, clover-late variant (depth 2, 6 RHS per CTA, 2 CTAs/SM) vs product
mkdir -p gpurun_out
for tag in "" cl; do
  export B200_LIB_TAG=$tag
  echo "== variant '${tag:-product}'"
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_rhs or qprop or symmetric or twisted" > gpurun_out/r02s_pytest_${tag:-product}.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/r02s_pytest_${tag:-product}.log
  PROF_LATT=48,48,48,96 PROF_REPS=5 timeout 300 python scripts/prof_mrhs.py 2>&1 | tail -2
  timeout 300 python bench.py --no-cpu --no-fp32 --no-solve --steps 10 --warmup 3 > gpurun_out/r02s_bench_${tag:-product}.json 2> gpurun_out/r02s_bench_${tag:-product}.err
  python -c "
import json;b=json.loads(open('gpurun_out/r02s_bench_${tag:-product}.json').read().strip().splitlines()[-1])['multi_rhs'];print('fp64 M %.3f ms frac %.3f  cg iter %.2f ms'%(b['clover_dslash']['ms_per_apply'],b['clover_dslash']['frac_of_peak'],b['ms_per_iteration']))"
done
