#!/bin/bash
# round-2 GPU job F (2 GPUs): full suite on the final build (new LDL^dagger kernel, strided pack CTAs, BLAS unroll, NT download copy),
# bench at N=1 and N=2 on the same box, 64^3x128 on one GPU with the z-chunked traversal, setup timing
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02f_pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_pytest_2gpu.log
tail -4 gpurun_out/r02f_pytest_2gpu.log
python bench.py > gpurun_out/r02f_bench_1gpu.json 2> gpurun_out/r02f_bench_1gpu.err; echo "bench1 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 2 \
  > gpurun_out/r02f_bench_2gpu.json 2> gpurun_out/r02f_bench_2gpu.err; echo "bench2 rc=$?"
python bench.py --lattice 64 64 64 128 --solver BICGSTAB --nrhs 1 --no-cpu --no-fp32 --steps 10 --warmup 3 > gpurun_out/r02f_bench_64x128_1gpu.json 2> gpurun_out/r02f_bench_64x128_1gpu.err; echo "bench 64 rc=$?"
python scripts/time_setup.py > gpurun_out/r02f_setup.json 2> gpurun_out/r02f_setup.err; echo "setup rc=$?"
tail -c 300 gpurun_out/r02f_bench_2gpu.err
