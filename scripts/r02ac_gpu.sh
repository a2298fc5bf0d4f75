#!/bin/bash
# round-2 GPU job AC (1 GPU): validation of the final build (batched split reductions, block-wise extra operand) -- pytest -m gpu, smoke, bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02ac_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ac_pytest.log
tail -4 gpurun_out/r02ac_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02ac_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02ac_smoke.log
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02ac_bench_1gpu.json 2> gpurun_out/r02ac_bench_1gpu.err ) 2>&1 | grep real; echo "bench rc=$?"
