#!/bin/bash
# round-2 GPU job O (1 GPU): L2 prefetch distance of the batched kernels
mkdir -p gpurun_out
B200_MRHS_PF_BLOCKS=7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_rhs or qprop or symmetric_operator" > gpurun_out/r02o_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02o_pytest.log
for pf in 0 150 300 600 1200 2400 4800; do
  B200_MRHS_PF_BLOCKS=$pf python bench.py --no-cpu --no-fp32 --no-solve --steps 10 --warmup 3 > gpurun_out/r02o_bench_pf$pf.json 2> gpurun_out/r02o_bench_pf$pf.err
  python -c "
import json;b=json.loads(open('gpurun_out/r02o_bench_pf$pf.json').read().strip().splitlines()[-1])['multi_rhs'];print('pf=$pf fp64 M %.3f ms frac %.3f  cg iter %.2f ms'%(b['clover_dslash']['ms_per_apply'],b['clover_dslash']['frac_of_peak'],b['ms_per_iteration']))"
done
for pf in 0 600 2400; do
  B200_MRHS_PF_BLOCKS=$pf python bench.py --prec single --no-cpu --no-solve --steps 10 --warmup 3 > gpurun_out/r02o_bench_f32_pf$pf.json 2> gpurun_out/r02o_bench_f32_pf$pf.err
  python -c "
import json;b=json.loads(open('gpurun_out/r02o_bench_f32_pf$pf.json').read().strip().splitlines()[-1])['multi_rhs'];print('pf=$pf fp32 M %.3f ms frac %.3f  cg iter %.2f ms'%(b['clover_dslash']['ms_per_apply'],b['clover_dslash']['frac_of_peak'],b['ms_per_iteration']))"
done
