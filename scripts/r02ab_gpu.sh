#!/bin/bash
# round-2 GPU job AB (1 GPU): split reductions in the batched kernels (dslash_mrhs_finish_kernel) -- parity, then the batched
# CG / BiCGStab iteration kernel by kernel
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_multi_gpu.py -m gpu -x -q -k "multi_rhs or qprop or split or one_device_split" > gpurun_out/r02ab_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02ab_pytest.log
python bench.py --no-cpu --no-fp32 --no-solve --steps 10 --warmup 3 > gpurun_out/r02ab_bench_cg.json 2> gpurun_out/r02ab_bench_cg.err; echo "rc=$?"
python bench.py --solver BICGSTAB --no-cpu --no-fp32 --no-solve --steps 10 --warmup 3 > gpurun_out/r02ab_bench_bicgstab.json 2> gpurun_out/r02ab_bench_bicgstab.err; echo "rc=$?"
python -c "
import json
for f in ['cg','bicgstab']:
    b=json.loads(open('gpurun_out/r02ab_bench_%s.json'%f).read().strip().splitlines()[-1])['multi_rhs']
    print(f, b['ms_per_iteration'])
    for k in b['kernels_in_loop']: print('   %-40s x%d %.3f ms frac %.3f share %.3f'%(k['kernel'],k['launches_per_iteration'],k['ms_per_launch'],k['frac'],k['share_of_iteration']))
"
