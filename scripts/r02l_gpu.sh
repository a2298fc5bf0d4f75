#!/bin/bash
# round-2 GPU job L (1 GPU): ncu launch list of the bench command and ncu --set full of the loop's Dslash kernels, FINAL build
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "recon12_rejects or zero_initial" > gpurun_out/r02l_pytest.log 2>&1; tail -3 gpurun_out/r02l_pytest.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --nrhs 1 --no-fp32 --no-solve > gpurun_out/r02l_launches.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none -k regex:"dslash_kernel|dslash_finish|cg_update" --launch-skip 20 -c 14 -f -o gpurun_out/r02l_full \
    python bench.py --no-cpu --nrhs 1 --no-solve --no-fp32 --steps 2 --warmup 1 > gpurun_out/r02l_ncu.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/r02l_full.ncu-rep --page raw --csv > gpurun_out/r02l_full_raw.csv 2>/dev/null
rm -f gpurun_out/r02l_full.ncu-rep
ls -la gpurun_out | tail -8
