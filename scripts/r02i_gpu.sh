#!/bin/bash
# round-2 GPU job I (1 GPU): the tile-sharing batched kernel -- parity tests, then the batched leg of the bench with and without it
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_rhs or qprop or symmetric_operator or twisted" > gpurun_out/r02i_pytest_mrhs.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02i_pytest_mrhs.log
tail -15 gpurun_out/r02i_pytest_mrhs.log
for tile in 1 0; do
  B200_MRHS_TILE=$tile python bench.py --no-cpu --no-fp32 --no-solve --steps 10 --warmup 3 > gpurun_out/r02i_bench_tile$tile.json 2> gpurun_out/r02i_bench_tile$tile.err; echo "bench tile=$tile rc=$?"
  python -c "
import json;b=json.loads(open('gpurun_out/r02i_bench_tile$tile.json').read().strip().splitlines()[-1]);print(json.dumps(b['multi_rhs']))"
done
B200_MRHS_TILE=1 python bench.py --prec single --no-cpu --no-solve --steps 10 --warmup 3 > gpurun_out/r02i_bench_f32_tile1.json 2> gpurun_out/r02i_bench_f32_tile1.err
B200_MRHS_TILE=0 python bench.py --prec single --no-cpu --no-solve --steps 10 --warmup 3 > gpurun_out/r02i_bench_f32_tile0.json 2> gpurun_out/r02i_bench_f32_tile0.err
for t in 1 0; do python -c "
import json;b=json.loads(open('gpurun_out/r02i_bench_f32_tile$t.json').read().strip().splitlines()[-1]);print('f32 tile=$t',json.dumps(b['multi_rhs']))"; done
