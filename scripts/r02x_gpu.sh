#!/bin/bash
# round-2 GPU job X (1 GPU): the batched CG / BiCGStab iteration kernel by kernel (bench.py multi_rhs.kernels_in_loop)
mkdir -p gpurun_out
python bench.py --no-cpu --no-fp32 --no-solve --steps 10 --warmup 3 > gpurun_out/r02x_bench_cg.json 2> gpurun_out/r02x_bench_cg.err; echo "rc=$?"
python bench.py --solver BICGSTAB --no-cpu --no-fp32 --no-solve --steps 10 --warmup 3 > gpurun_out/r02x_bench_bicgstab.json 2> gpurun_out/r02x_bench_bicgstab.err; echo "rc=$?"
python -c "
import json
for f in ['cg','bicgstab']:
    b=json.loads(open('gpurun_out/r02x_bench_%s.json'%f).read().strip().splitlines()[-1])['multi_rhs']
    print(f, b['ms_per_iteration'])
    for k in b['kernels_in_loop']: print('   %-40s x%d %.3f ms frac %.3f share %.3f'%(k['kernel'],k['launches_per_iteration'],k['ms_per_launch'],k['frac'],k['share_of_iteration']))
"
