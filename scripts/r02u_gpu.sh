#!/bin/bash
# round-2 GPU job U (8 GPUs): multi-GPU parity (2 / 4 / 8 GPUs, batched fields on split lattices included) and BASELINE config 5
# (64^3x128 BiCGStab, 12 right-hand sides, Z x T = 2 x 4) with the rewritten batched kernel
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python -m pytest tests/test_multi_gpu.py -m gpu -x -q -k "two_gpu_parity or eight_gpu or txz_grid" > gpurun_out/r02u_pytest_mgpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02u_pytest_mgpu.log
tail -4 gpurun_out/r02u_pytest_mgpu.log
$TR --nproc-per-node 8 --master-port 29612 bench.py --gpus 8 --lattice 64 64 64 128 --grid 2 4 --solver BICGSTAB --nrhs 12 --no-cpu \
   > gpurun_out/r02u_bench_8gpu_config5.json 2> gpurun_out/r02u_bench_8gpu_config5.err; echo "bench8 config5 rc=$?"
tail -c 300 gpurun_out/r02u_bench_8gpu_config5.err
