#!/bin/bash
# round-2 GPU job AA (1 GPU): z-chunk thickness of the batched traversal (B200_MRHS_L2_KB -> d z-planes per chunk at 48^3 x 12 RHS)
mkdir -p gpurun_out
for kb in 16500 24500 32700 49000 65300 98000 200000; do
  B200_MRHS_L2_KB=$kb python bench.py --no-cpu --no-fp32 --no-solve --steps 6 --warmup 2 > gpurun_out/r02aa_bench_$kb.json 2> gpurun_out/r02aa_bench_$kb.err
  python -c "
import json
b=json.loads(open('gpurun_out/r02aa_bench_$kb.json').read().strip().splitlines()[-1])['multi_rhs']
print('L2_KB=$kb iter %.2f ms  M(AINV+M) %.2f ms frac %.3f | '%(b['ms_per_iteration'],b['clover_dslash']['ms_per_apply'],b['clover_dslash']['frac_of_peak']) + ' '.join('%.2f'%k['ms_per_launch'] for k in b['kernels_in_loop']))"
done
