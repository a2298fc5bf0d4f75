#!/bin/bash
# round-2 GPU job G (8 GPUs): final build (LL allreduce, strided pack CTAs): 8/4-GPU parity cases, bench at N=8, 4, 2, 1 on the SAME box
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python -m pytest tests/test_multi_gpu.py -m gpu -x -q -k "eight_gpu or four_gpu or two_gpu_thin" > gpurun_out/r02g_pytest_8gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02g_pytest_8gpu.log
tail -4 gpurun_out/r02g_pytest_8gpu.log
$TR --nproc-per-node 8 --master-port 29611 bench.py --gpus 8 > gpurun_out/r02g_bench_8gpu.json 2> gpurun_out/r02g_bench_8gpu.err; echo "bench8 rc=$?"
$TR --nproc-per-node 4 --master-port 29613 bench.py --gpus 4 --no-fp32 > gpurun_out/r02g_bench_4gpu.json 2> gpurun_out/r02g_bench_4gpu.err; echo "bench4 rc=$?"
$TR --nproc-per-node 2 --master-port 29614 bench.py --gpus 2 --no-fp32 > gpurun_out/r02g_bench_2gpu.json 2> gpurun_out/r02g_bench_2gpu.err; echo "bench2 rc=$?"
python bench.py --no-fp32 --no-cpu > gpurun_out/r02g_bench_1gpu.json 2> gpurun_out/r02g_bench_1gpu.err; echo "bench1 rc=$?"
$TR --nproc-per-node 8 --master-port 29612 bench.py --gpus 8 --lattice 64 64 64 128 --grid 2 4 --solver BICGSTAB --nrhs 12 --no-cpu \
   > gpurun_out/r02g_bench_8gpu_config5.json 2> gpurun_out/r02g_bench_8gpu_config5.err; echo "bench8 config5 rc=$?"
for nt in 1 0; do B200_NT_COPY=$nt python scripts/time_setup.py > gpurun_out/r02g_setup_nt$nt.json 2> gpurun_out/r02g_setup_nt$nt.err; done
tail -c 300 gpurun_out/r02g_bench_8gpu.err
