#!/bin/bash
# round-2 GPU job C (2 GPUs): full GPU suite incl. the real 2-GPU cases, bench at N=1 and N=2 on the same box
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02c_pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_pytest_2gpu.log
tail -4 gpurun_out/r02c_pytest_2gpu.log
python bench.py > gpurun_out/r02c_bench_1gpu.json 2> gpurun_out/r02c_bench_1gpu.err; echo "bench1 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 2 \
  > gpurun_out/r02c_bench_2gpu.json 2> gpurun_out/r02c_bench_2gpu.err; echo "bench2 rc=$?"
tail -c 400 gpurun_out/r02c_bench_2gpu.err
