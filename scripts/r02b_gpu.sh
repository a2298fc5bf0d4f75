#!/bin/bash
# round-2 GPU job B: golden GPU tests, A/B of the reduction tail variants, ncu --set full of the loop's Dslash kernels
mkdir -p gpurun_out
python -m pytest tests/test_golden.py tests/test_adapter_exec.py -m gpu -x -q > gpurun_out/r02b_pytest_golden.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest_golden.log
tail -3 gpurun_out/r02b_pytest_golden.log
for tag in default probe1 probe2 oldfence; do
  if [ "$tag" = default ]; then unset B200_LIB_TAG; else export B200_LIB_TAG=$tag; fi
  python bench.py --no-cpu --nrhs 1 --no-solve --no-fp32 --steps 10 --warmup 3 > gpurun_out/r02b_bench_$tag.json 2> gpurun_out/r02b_bench_$tag.err
  echo "bench $tag rc=$?"
done
unset B200_LIB_TAG
ncu --set full --clock-control none --import-source on -k regex:dslash_kernel -c 8 -f -o gpurun_out/r02b_dslash \
    python bench.py --no-cpu --nrhs 1 --no-solve --no-fp32 --steps 2 --warmup 1 > gpurun_out/r02b_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/r02b_dslash.ncu-rep --page raw --csv > gpurun_out/r02b_dslash_raw.csv 2>/dev/null
ncu -i gpurun_out/r02b_dslash.ncu-rep --page source --csv --print-source sass > gpurun_out/r02b_dslash_source.csv 2>/dev/null
rm -f gpurun_out/r02b_dslash.ncu-rep     # 77 MB: gpurun only brings back 64 MiB; the two CSV exports carry what is read
gzip -f gpurun_out/r02b_dslash_source.csv
ls -la gpurun_out/ | tail -20
