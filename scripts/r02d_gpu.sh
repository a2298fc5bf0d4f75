#!/bin/bash
# round-2 GPU job D (1 GPU): full suite with the rebuilt library, 64^3x128 N=1 record, compute-sanitizer runs of the adapter harness
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_pytest.log
tail -4 gpurun_out/r02d_pytest.log
python bench.py --lattice 64 64 64 128 --solver BICGSTAB --nrhs 1 --no-cpu --no-fp32 --steps 10 --warmup 3 > gpurun_out/r02d_bench_64x128_1gpu.json 2> gpurun_out/r02d_bench_64x128_1gpu.err; echo "bench 64 rc=$?"
python bench.py --solver BICGSTAB --nrhs 1 --no-cpu --no-fp32 --steps 10 --warmup 3 > gpurun_out/r02d_bench_bicgstab_1gpu.json 2> gpurun_out/r02d_bench_bicgstab_1gpu.err; echo "bench bicg rc=$?"
bash scripts/build_adapter_exec.sh gpurun_out/adapter_exec > /dev/null
for tool in memcheck racecheck synccheck; do
  OMP_NUM_THREADS=2 timeout 600 compute-sanitizer --tool $tool --target-processes all gpurun_out/adapter_exec 4 4 4 8 1 > gpurun_out/r02d_sanitizer_${tool}_1rank.log 2>&1; echo "$tool 1rank rc=$?"
done
OMP_NUM_THREADS=2 B200_PEER_TIMEOUT_S=600 timeout 900 compute-sanitizer --tool memcheck --target-processes all gpurun_out/adapter_exec 4 4 4 8 2 > gpurun_out/r02d_sanitizer_memcheck_2rank.log 2>&1; echo "memcheck 2rank rc=$?"
OMP_NUM_THREADS=2 B200_PEER_TIMEOUT_S=600 timeout 900 compute-sanitizer --tool racecheck --target-processes all gpurun_out/adapter_exec 4 4 4 8 2 > gpurun_out/r02d_sanitizer_racecheck_2rank.log 2>&1; echo "racecheck 2rank rc=$?"
rm -f gpurun_out/adapter_exec
for f in gpurun_out/r02d_sanitizer_*.log; do echo "== $f"; tail -3 $f; done
