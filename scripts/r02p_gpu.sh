#!/bin/bash
# round-2 GPU job P (1 GPU): batched kernel with the lean prologue (fast site decode, grouped staging, early first fetch)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_rhs or qprop or symmetric or twisted or mdagm" > gpurun_out/r02p_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02p_pytest.log
PROF_LATT=48,48,48,96 PROF_REPS=5 python scripts/prof_mrhs.py 2>&1 | tail -2
PROF_LATT=48,48,48,96 PROF_REPS=5 PROF_PREC=single python scripts/prof_mrhs.py 2>&1 | tail -2
python bench.py --no-cpu --no-fp32 --no-solve --steps 10 --warmup 3 > gpurun_out/r02p_bench.json 2> gpurun_out/r02p_bench.err
python -c "
import json;b=json.loads(open('gpurun_out/r02p_bench.json').read().strip().splitlines()[-1])['multi_rhs'];print('fp64 M %.3f ms frac %.3f  cg iter %.2f ms'%(b['clover_dslash']['ms_per_apply'],b['clover_dslash']['frac_of_peak'],b['ms_per_iteration']))"
python bench.py --prec single --no-cpu --no-solve --steps 10 --warmup 3 > gpurun_out/r02p_bench_f32.json 2> gpurun_out/r02p_bench_f32.err
python -c "
import json;b=json.loads(open('gpurun_out/r02p_bench_f32.json').read().strip().splitlines()[-1])['multi_rhs'];print('fp32 M %.3f ms frac %.3f  cg iter %.2f ms'%(b['clover_dslash']['ms_per_apply'],b['clover_dslash']['frac_of_peak'],b['ms_per_iteration']))"
