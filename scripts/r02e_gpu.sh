#!/bin/bash
# round-2 GPU job E (8 GPUs): the never-executed 8-GPU parity cases, bench at N=8 (48^3x96 CG, T split) and BASELINE config 5
# (64^3x128 BiCGStab, 12 right-hand sides, Z x T = 2 x 4), bench at N=4 on the same box
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python -m pytest tests/test_multi_gpu.py -m gpu -x -q -k "eight_gpu or four_gpu" > gpurun_out/r02e_pytest_8gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02e_pytest_8gpu.log
tail -4 gpurun_out/r02e_pytest_8gpu.log
$TR --nproc-per-node 8 --master-port 29611 bench.py --gpus 8 > gpurun_out/r02e_bench_8gpu.json 2> gpurun_out/r02e_bench_8gpu.err; echo "bench8 rc=$?"
$TR --nproc-per-node 8 --master-port 29612 bench.py --gpus 8 --lattice 64 64 64 128 --grid 2 4 --solver BICGSTAB --nrhs 12 --no-cpu \
   > gpurun_out/r02e_bench_8gpu_config5.json 2> gpurun_out/r02e_bench_8gpu_config5.err; echo "bench8 config5 rc=$?"
$TR --nproc-per-node 4 --master-port 29613 bench.py --gpus 4 > gpurun_out/r02e_bench_4gpu.json 2> gpurun_out/r02e_bench_4gpu.err; echo "bench4 rc=$?"
tail -c 300 gpurun_out/r02e_bench_8gpu.err; tail -c 300 gpurun_out/r02e_bench_8gpu_config5.err
