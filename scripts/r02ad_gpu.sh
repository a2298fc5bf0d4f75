#!/bin/bash
# round-2 GPU job AD (2 GPUs): 2-GPU parity cases and the N=2 bench line (T split) with the final build
mkdir -p gpurun_out
python -m pytest tests/test_multi_gpu.py -m gpu -x -q -k "two_gpu" > gpurun_out/r02ad_pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ad_pytest_2gpu.log
tail -3 gpurun_out/r02ad_pytest_2gpu.log
python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2 --master-port 29621 bench.py --gpus 2 --no-cpu --no-fp32 > gpurun_out/r02ad_bench_2gpu.json 2> gpurun_out/r02ad_bench_2gpu.err; echo "bench2 rc=$?"
python -c "
import json
b=json.loads(open('gpurun_out/r02ad_bench_2gpu.json').read().strip().splitlines()[-1])
print('N=2 ms/step %.3f value %.0f solve %s it %s'%(b['ms_per_step'],b['value'],b['solve']['seconds'],b['solve']['iterations']), b['solve'].get('check'))
m=b.get('multi_rhs'); print('mrhs', m and m.get('ms_per_iteration'))
"
