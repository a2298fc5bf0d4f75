#!/bin/bash
# round-2 GPU job M (1 GPU): row-streaming link multiply in the batched kernels -- parity, then A/B of CTA shapes
mkdir -p gpurun_out
for tag in default rs8; do
  if [ "$tag" = default ]; then unset B200_LIB_TAG; else export B200_LIB_TAG=$tag; fi
  python -m pytest tests/test_gpu_parity.py tests/test_multi_gpu.py -m gpu -x -q -k "multi_rhs or qprop or symmetric_operator or twisted or one_device_grid_parity" > gpurun_out/r02m_pytest_$tag.log 2>&1; echo "pytest $tag rc=$?"; tail -2 gpurun_out/r02m_pytest_$tag.log
done
for tag in default rs8 base; do
  if [ "$tag" = default ]; then unset B200_LIB_TAG; else export B200_LIB_TAG=$tag; fi
  python bench.py --no-cpu --no-fp32 --no-solve --steps 10 --warmup 3 > gpurun_out/r02m_bench_$tag.json 2> gpurun_out/r02m_bench_$tag.err; echo "bench $tag rc=$?"
  python -c "
import json;b=json.loads(open('gpurun_out/r02m_bench_$tag.json').read().strip().splitlines()[-1]);print('$tag fp64',json.dumps(b['multi_rhs']))"
  python bench.py --prec single --no-cpu --no-solve --steps 10 --warmup 3 > gpurun_out/r02m_bench_f32_$tag.json 2> gpurun_out/r02m_bench_f32_$tag.err
  python -c "
import json;b=json.loads(open('gpurun_out/r02m_bench_f32_$tag.json').read().strip().splitlines()[-1]);print('$tag fp32',json.dumps(b['multi_rhs']))"
done
