#!/bin/bash
# round-2 GPU job AF (1 GPU): (z, y) chunks of the batched traversal against z-only chunks (B200_MRHS_YCHUNK=0)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_rhs or qprop_full or split_reduction" > gpurun_out/r02af_pytest.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/r02af_pytest.log
for y in 1 0; do
  B200_MRHS_YCHUNK=$y python bench.py --no-cpu --no-fp32 --no-solve --steps 8 --warmup 2 > gpurun_out/r02af_bench_y$y.json 2> gpurun_out/r02af_bench_y$y.err
  python -c "
import json
b=json.loads(open('gpurun_out/r02af_bench_y$y.json').read().strip().splitlines()[-1])['multi_rhs']
print('YCHUNK=$y iter %.2f ms  M(AINV+M) %.2f ms frac %.3f | '%(b['ms_per_iteration'],b['clover_dslash']['ms_per_apply'],b['clover_dslash']['frac_of_peak']) + ' '.join('%.2f'%k['ms_per_launch'] for k in b['kernels_in_loop']))"
done
